// scar_rawdata.cu — the body of ScarSearch.raw_data_output (reference scarmapper/INDEL_Processing.py:702-757)
// written natively from the per-read records: one line per scarred read of a sample,
//   left deletions, right deletions, deletion size, microhomology, insertion, insertion size,
//   consensus left / right junction, reference left / right junction, consensus, target region
// with the swaps and reverse complements the reference applies for 3'-strand guides (:733-746) and its
// "skip unaltered reads" rule (:748-750).  Host code (formatting and file output are the host's job);
// the reference builds the same text by Python string concatenation over a per-read list of lists.
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>

#include "scar_internal.cuh"
#include "scar_comp.h"

int scar_fail(int code, const char *fmt, ...);

namespace {
inline uint8_t comp(uint8_t c) { return scar_comp(c); }

struct Sink {                       // counts, and writes when there is a buffer
    uint8_t *out; int64_t cap, n;
    void put(const uint8_t *p, int64_t len) { if (out && n + len <= cap) memcpy(out + n, p, (size_t)len); n += len; }
    void put(char c) { if (out && n + 1 <= cap) out[n] = (uint8_t)c; ++n; }
    void num(long long v) { char b[24]; const int k = snprintf(b, sizeof b, "%lld", v); put((const uint8_t *)b, k); }
    // s[a:b) forward, or its reverse complement
    void seq(const std::string &s, int a, int b, bool rc)
    {
        if (b <= a) return;
        if (!rc) { put((const uint8_t *)s.data() + a, b - a); return; }
        for (int k = b - 1; k >= a; --k) put((char)comp((uint8_t)s[(size_t)k]));
    }
};
}   // namespace

// the sequence ScarSearch is handed for a raw read (INDEL_Processing.py:946,978,981)
static void stored_sequence(const uint8_t *raw, int rawlen, int32_t platform, std::string &stored)
{
    if (platform == SCAR_PLATFORM_TRUSEQ) stored.assign((const char *)raw + (rawlen > 6 ? 6 : rawlen), (size_t)(rawlen > 12 ? rawlen - 12 : 0));
    else if (platform == SCAR_PLATFORM_RAMSDEN) {
        stored.resize((size_t)rawlen);
        for (int k = 0; k < rawlen; ++k) stored[(size_t)k] = (char)comp(raw[rawlen - 1 - k]);
    } else stored.assign((const char *)raw, (size_t)rawlen);
}

// one Raw_Data line of a scarred read (nothing for an unaltered read, :748-750)
static void raw_row(Sink &S, const scar_record &R, const std::string &stored, const std::string &region, int32_t target_len,
                    int32_t cutsite, bool rc)
{
    const int len = (int)stored.size();
    const int clj = R.clj, crj = R.crj, tlj = R.tlj, trj = R.trj;
    const int ldel = cutsite > tlj ? cutsite - tlj : 0;                       // len(target[tlj:cutsite])
    const int rdel = (trj < target_len ? trj : target_len) > cutsite ? (trj < target_len ? trj : target_len) - cutsite : 0;
    int ia = 0, ib = 0, ma = 0, mb = 0;                                         // insertion / microhomology slices
    if (R.flags & SCAR_FL_INSERTION) { ia = clj < len ? clj : len; ib = crj < len ? crj : len; if (ib < ia) ib = ia; }
    if (R.flags & SCAR_FL_MICROHOM) { ma = crj < len ? crj : len; mb = clj < len ? clj : len; if (mb < ma) mb = ma; }
    const int ins_size = ib - ia, mh_size = mb - ma;
    const int del_size = ldel + rdel + mh_size;
    if (del_size == 0 && ins_size == 0) return;                                 // skip unaltered reads (:748-750)
    S.num(rc ? rdel : ldel); S.put('\t');
    S.num(rc ? ldel : rdel); S.put('\t');
    S.num(del_size); S.put('\t');
    S.seq(stored, ma, mb, rc); S.put('\t');
    S.seq(stored, ia, ib, rc); S.put('\t');
    S.num(ins_size); S.put('\t');
    S.num(rc ? len - crj : clj); S.put('\t');
    S.num(rc ? len - clj : crj); S.put('\t');
    S.num(rc ? target_len - trj : tlj); S.put('\t');
    S.num(rc ? target_len - tlj : trj); S.put('\t');
    S.seq(stored, 0, len, rc); S.put('\t');
    S.seq(region, 0, target_len, rc); S.put('\n');
}

extern "C" int scar_raw_rows_host(const scar_record *records, const scar_ragged *reads, int64_t n, int32_t sample,
                                  int32_t platform, const uint8_t *target_region, int32_t target_len, int32_t cutsite,
                                  int32_t reverse_comp, uint8_t *out, int64_t out_capacity, int64_t *out_bytes)
{
    if (!records || !reads || !target_region || !out_bytes || n < 0 || target_len < 0)
        return scar_fail(SCAR_ERR_INVALID, "bad argument");
    Sink S{out, out ? out_capacity : 0, 0};
    const std::string region((const char *)target_region, (size_t)target_len);
    std::string stored;
    for (int64_t r = 0; r < n; ++r) {
        const scar_record &R = records[r];
        if (R.status != SCAR_ST_SCAR || (int32_t)R.sample != sample) continue;
        const uint8_t *raw; int rawlen;
        get_item(*reads, r, raw, rawlen);
        stored_sequence(raw, rawlen, platform, stored);
        raw_row(S, R, stored, region, target_len, cutsite, reverse_comp != 0);
    }
    *out_bytes = S.n;
    if (out && S.n > out_capacity) return scar_fail(SCAR_ERR_INVALID, "raw rows: %lld bytes needed", (long long)S.n);
    return SCAR_OK;
}

// The same for a LIST of reads (rows[k] = read index, the scarred reads of one sample in read order): a run with
// many samples buckets the reads once and pays O(reads of the sample) per file instead of a scan of the whole run.
extern "C" int scar_raw_rows_indexed_host(const scar_record *records, const scar_ragged *reads, const int64_t *rows,
                                          int64_t n_rows, int32_t platform, const uint8_t *target_region,
                                          int32_t target_len, int32_t cutsite, int32_t reverse_comp, uint8_t *out,
                                          int64_t out_capacity, int64_t *out_bytes)
{
    if (!records || !reads || (!rows && n_rows) || !target_region || !out_bytes || n_rows < 0 || target_len < 0)
        return scar_fail(SCAR_ERR_INVALID, "bad argument");
    Sink S{out, out ? out_capacity : 0, 0};
    const std::string region((const char *)target_region, (size_t)target_len);
    std::string stored;
    for (int64_t k = 0; k < n_rows; ++k) {
        const int64_t r = rows[k];
        const scar_record &R = records[r];
        if (R.status != SCAR_ST_SCAR) continue;
        const uint8_t *raw; int rawlen;
        get_item(*reads, r, raw, rawlen);
        stored_sequence(raw, rawlen, platform, stored);
        raw_row(S, R, stored, region, target_len, cutsite, reverse_comp != 0);
    }
    *out_bytes = S.n;
    if (out && S.n > out_capacity) return scar_fail(SCAR_ERR_INVALID, "raw rows: %lld bytes needed", (long long)S.n);
    return SCAR_OK;
}

// Stored sequences (platform transform applied) of a list of reads, gathered into one CSR buffer; row k is reverse-
// complemented when flip[k] != 0 (3'-strand guides: frequency_output works on rcomp(consensus), :312-327).
// out_off receives n_rows + 1 offsets; call with out = NULL for the size.
namespace {
struct CompTable { uint8_t t[256]; CompTable() { for (int i = 0; i < 256; ++i) t[i] = comp((uint8_t)i); } };
const CompTable COMP_TABLE;

// f(t, lo, hi) over [0, n) on up to 8 threads (row blocks); small jobs stay on the calling thread
template <class F> void for_row_blocks(int64_t n, F f)
{
    const int threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(8, (int64_t)std::thread::hardware_concurrency()), n / 32768));
    if (threads <= 1) { f(0, n); return; }
    std::vector<std::thread> th;
    const int64_t per = (n + threads - 1) / threads;
    for (int t = 1; t < threads; ++t) th.emplace_back([=] { f(std::min(n, t * per), std::min(n, (t + 1) * per)); });
    f(0, std::min(n, per));
    for (auto &x : th) x.join();
}
}   // namespace

extern "C" int scar_stored_gather_host(const scar_ragged *reads, const int64_t *rows, int64_t n_rows, int32_t platform,
                                       const uint8_t *flip, uint8_t *out, int64_t out_capacity, int64_t *out_off)
{
    if (!reads || (!rows && n_rows) || !out_off || n_rows < 0) return scar_fail(SCAR_ERR_INVALID, "bad argument");
    // the stored sequence is a span of the raw read, forwards or as its reverse complement (Ramsden stores rcomp; a
    // flipped row is complemented once more - comp is an involution, so the two cancel)
    auto span = [&](int64_t k, const uint8_t *&p, int &len) {
        const uint8_t *raw; int rawlen;
        get_item(*reads, rows[k], raw, rawlen);
        if (platform == SCAR_PLATFORM_TRUSEQ) { p = raw + (rawlen > 6 ? 6 : rawlen); len = rawlen > 12 ? rawlen - 12 : 0; }
        else { p = raw; len = rawlen; }
    };
    out_off[0] = 0;
    for_row_blocks(n_rows, [&](int64_t lo, int64_t hi) {
        for (int64_t k = lo; k < hi; ++k) { const uint8_t *p; int len; span(k, p, len); out_off[k + 1] = len; }
    });
    for (int64_t k = 0; k < n_rows; ++k) out_off[k + 1] += out_off[k];
    const int64_t total = out_off[n_rows];
    if (!out) return SCAR_OK;
    if (total > out_capacity) return scar_fail(SCAR_ERR_INVALID, "stored gather: %lld bytes needed", (long long)total);
    const bool ramsden = platform == SCAR_PLATFORM_RAMSDEN;
    for_row_blocks(n_rows, [&](int64_t lo, int64_t hi) {
        for (int64_t k = lo; k < hi; ++k) {
            const uint8_t *p; int len;
            span(k, p, len);
            uint8_t *dst = out + out_off[k];
            if (ramsden != (flip && flip[k])) for (int i = 0; i < len; ++i) dst[i] = COMP_TABLE.t[p[len - 1 - i]];
            else memcpy(dst, p, (size_t)len);
        }
    });
    return SCAR_OK;
}

// ---- demultiplexed FASTQ (--Demultiplex True) -------------------------------------------------------------------
// The reference keeps [name, seq, qual] of every read per sample and writes "@name\nseq\n+\nqual\n" per read at the
// end of the run (scarmapper/INDEL_Processing.py:876-889,988-1002; Valkyries/FASTQ_Tools.py:532-551), with
// name = header.strip("\n").strip("@") and seq / qual stripped of white space (FASTQ_Tools.py:640-643).  Here the
// records are formatted straight from the FASTQ text: rows[k] is a read index, seq_off[r] the offset of read r's
// (stripped) sequence in `text` as scar_fastq_index reports it; the header is the line before the sequence line, the
// quality line the second line after it.  Call with out = NULL for the size.
static inline bool fq_space(uint8_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f' || c == '\n'; }

extern "C" int scar_demux_fastq_rows_host(const uint8_t *text, int64_t n_bytes, const int64_t *seq_off, const int32_t *seq_len,
                                          const int64_t *rows, int64_t n_rows, uint8_t *out, int64_t out_capacity,
                                          int64_t *out_bytes)
{
    if (!text || !seq_off || !seq_len || (!rows && n_rows) || !out_bytes || n_rows < 0 || n_bytes < 0)
        return scar_fail(SCAR_ERR_INVALID, "bad argument");
    Sink S{out, out ? out_capacity : 0, 0};
    for (int64_t k = 0; k < n_rows; ++k) {
        const int64_t r = rows[k];
        const int64_t s0 = seq_off[r], s1 = s0 + seq_len[r];
        if (s0 < 0 || s1 > n_bytes) return scar_fail(SCAR_ERR_INVALID, "row %lld: sequence view outside the text", (long long)k);
        // header line: [h0, h1) is the line that ends just before the sequence line starts
        int64_t ls = s0;
        while (ls > 0 && text[ls - 1] != '\n') --ls;                 // start of the sequence line
        if (ls == 0) return scar_fail(SCAR_ERR_INVALID, "row %lld: no header line before the sequence", (long long)k);
        int64_t h1 = ls - 1, h0 = h1;                                // h1 = position of the header's newline
        while (h0 > 0 && text[h0 - 1] != '\n') --h0;
        if (h1 > h0 && text[h1 - 1] == '\r') --h1;                   // text mode: "\r\n" is one newline
        while (h0 < h1 && text[h0] == '@') ++h0;                     // .strip("@")
        while (h1 > h0 && text[h1 - 1] == '@') --h1;
        // quality line: skip the rest of the sequence line and the '+' line
        int64_t p = s1;
        while (p < n_bytes && text[p] != '\n') ++p;
        ++p;
        while (p < n_bytes && text[p] != '\n') ++p;
        int64_t q0 = p + 1, q1 = q0;
        if (q0 > n_bytes) q0 = q1 = n_bytes;
        while (q1 < n_bytes && text[q1] != '\n') ++q1;
        while (q0 < q1 && fq_space(text[q0])) ++q0;
        while (q1 > q0 && fq_space(text[q1 - 1])) --q1;
        S.put('@'); S.put(text + h0, h1 - h0); S.put('\n');
        S.put(text + s0, s1 - s0); S.put((const uint8_t *)"\n+\n", 3);
        S.put(text + q0, q1 - q0); S.put('\n');
    }
    *out_bytes = S.n;
    if (out && S.n > out_capacity) return scar_fail(SCAR_ERR_INVALID, "demultiplexed FASTQ: %lld bytes needed", (long long)S.n);
    return SCAR_OK;
}
