// scar_fastq.cu — host-side FASTQ record indexer (the batched loader's front end).
//
// Replaces FASTQ_Reader.seq_read (reference Valkyries/FASTQ_Tools.py:623-653: four lines per record,
// name = line1.strip("\n").strip("@"), seq/qual = line.strip("\n").strip(), ValueError when
// len(seq) != len(qual)) and the header split of index_matching
// (scarmapper/INDEL_Processing.py:1155-1156: name.split(":")[-1].split("+")[0..1]) for a whole text
// buffer: it emits (offset, length) VIEWS into the text for the sequence and the two index fields, so
// the text is copied to the GPU once and K1/K2 read the fields in place (scar_ragged views mode).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/scar_b200.h"

int scar_fail(int code, const char *fmt, ...);

static inline bool is_space(uint8_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f' || c == '\n'; }

extern "C" int scar_fastq_index(const uint8_t *text, int64_t n_bytes, int32_t is_final, int32_t want_index_fields,
                                int64_t capacity, int64_t *seq_off, int32_t *seq_len, int64_t *f0_off, int32_t *f0_len,
                                int64_t *f1_off, int32_t *f1_len, int64_t *n_records, int64_t *consumed)
{
    if (!text || n_bytes < 0 || !seq_off || !seq_len || !n_records || !consumed)
        return scar_fail(SCAR_ERR_INVALID, "bad argument");
    if (want_index_fields && (!f0_off || !f0_len || !f1_off || !f1_len))
        return scar_fail(SCAR_ERR_INVALID, "index field outputs missing");
    int64_t pos = 0, n = 0;
    while (pos < n_bytes && n < capacity) {
        int64_t ls[4], le[4];       // line starts / ends (end excludes the newline)
        int64_t p = pos;
        int k = 0;
        for (; k < 4; ++k) {
            if (p >= n_bytes) break;
            const uint8_t *nl = (const uint8_t *)memchr(text + p, '\n', (size_t)(n_bytes - p));
            ls[k] = p;
            if (nl) { le[k] = nl - text; p = le[k] + 1; }
            else if (is_final && k == 3) { le[k] = n_bytes; p = n_bytes; }   // last record without a newline
            else break;
        }
        if (k < 4) break;           // incomplete record: the caller feeds it again with more text
        // sequence and quality: surrounding whitespace stripped
        int64_t s0 = ls[1], s1 = le[1], q0 = ls[3], q1 = le[3];
        while (s0 < s1 && is_space(text[s0])) ++s0;
        while (s1 > s0 && is_space(text[s1 - 1])) --s1;
        while (q0 < q1 && is_space(text[q0])) ++q0;
        while (q1 > q0 && is_space(text[q1 - 1])) --q1;
        if (s1 - s0 != q1 - q0)
            return scar_fail(SCAR_ERR_INVALID, "record %lld: sequence and quality scores of different lengths (%lld vs %lld)",
                             (long long)n, (long long)(s1 - s0), (long long)(q1 - q0));
        if (s1 - s0 > 0x7FFFFFFF) return scar_fail(SCAR_ERR_INVALID, "record %lld: line too long", (long long)n);
        seq_off[n] = s0; seq_len[n] = (int32_t)(s1 - s0);
        if (want_index_fields) {
            int64_t h0 = ls[0], h1 = le[0];
            if (h1 > h0 && text[h1 - 1] == '\r') --h1;        // text mode: "\r\n" is one newline
            while (h0 < h1 && text[h0] == '@') ++h0;          // .strip("@")
            while (h1 > h0 && text[h1 - 1] == '@') --h1;
            int64_t c = h1;                                   // text after the last ':'
            while (c > h0 && text[c - 1] != ':') --c;
            const uint8_t *plus = (const uint8_t *)memchr(text + c, '+', (size_t)(h1 - c));
            if (!plus)
                return scar_fail(SCAR_ERR_INVALID, "record %lld: header has no '+' between the index reads "
                                 "(the reference raises IndexError)", (long long)n);
            const int64_t p1 = plus - text;
            const uint8_t *plus2 = (const uint8_t *)memchr(text + p1 + 1, '+', (size_t)(h1 - p1 - 1));
            const int64_t p2 = plus2 ? plus2 - text : h1;
            f0_off[n] = c; f0_len[n] = (int32_t)(p1 - c);
            f1_off[n] = p1 + 1; f1_len[n] = (int32_t)(p2 - p1 - 1);
        }
        ++n;
        pos = p;
    }
    *n_records = n;
    *consumed = pos;
    return SCAR_OK;
}
