// scar_pgzip.h — ONE gzip / DEFLATE stream inflated by several threads (host side of F2).
//
// Replaces, for the batched drop-in, what the reference does through Python's gzip module in FASTQ_Reader
// (Valkyries/FASTQ_Tools.py:562-653: `gzip.open(path, 'rt')`, one zlib stream, one core).  A DEFLATE stream has no
// index: a back-reference may reach 32 KiB into text produced by everything before it.  The scheme here is the
// two-pass one published as pugz (Kerbiriou & Chikhi 2019) / rapidgzip, written from that description:
//
//   pass 1  the compressed bytes are cut into chunks.  Every chunk but the first looks for the first bit position in
//           its range at which a non-final dynamic-Huffman block starts (header parses, the code-length code and both
//           alphabets are complete prefix codes, the block decodes to its end-of-block symbol and yields text bytes
//           only) and decodes from there with an UNKNOWN window: output symbols are 16 bit, a back-reference that
//           reaches in front of the chunk yields the marker 256 + (position in the 32 KiB window).  A chunk stops at the
//           block boundary where the next chunk started.
//   pass 2  in stream order: the last 32 KiB of a chunk, resolved with its own window, are the next chunk's window;
//           with its window known a chunk's symbols become bytes (markers looked up), piece by piece and side by side.
//
// Correctness never rests on the guess: chunk 0 starts at the true first block; a chunk is accepted only if the chunk
// before it ENDED on exactly the bit where it started (otherwise it is decoded again from there), and every member's
// CRC-32 and length are checked against its trailer (per-piece CRCs folded with crc32_combine).
// Host code only (no CUDA): tests/test_pgzip.py drives it on the CPU through scar_gunzip_host.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <sys/mman.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <zlib.h>          // crc32 / crc32_combine only; the inflate below is this file's own

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace scar_pgzip {

constexpr int WINDOW = 32768;
constexpr int LIT_PRIM = 10, DIST_PRIM = 8;
constexpr int LIT_TABLE = (1 << LIT_PRIM) + 32 * 288, DIST_TABLE = (1 << DIST_PRIM) + 128 * 32;

// op: 0 literal | 16+e length / distance base with e extra bits | 32 end of block | 64+s link to a sub-table of s bits
//     | 128 not a code
struct Entry { uint16_t val; uint8_t bits; uint8_t op; };

struct Bits {
    const uint8_t *p; int64_t n;       // the whole compressed input
    int64_t pos;                       // next byte to load (may run past n: zeros, caught by overrun())
    uint64_t buf; int cnt;             // cnt valid bits, least significant first
    void init(const uint8_t *in, int64_t len, uint64_t bit)
    {
        p = in; n = len; pos = (int64_t)(bit >> 3); buf = 0; cnt = 0;
        refill();
        drop((int)(bit & 7));
    }
    inline void refill()               // at least 56 bits afterwards
    {
        if (pos + 8 <= n) {
            uint64_t w;
            memcpy(&w, p + pos, 8);
            buf |= w << cnt;
            pos += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56) { buf |= (uint64_t)(pos < n ? p[pos] : 0) << cnt; ++pos; cnt += 8; }
        }
    }
    inline uint32_t peek(int k) const { return (uint32_t)(buf & ((1ull << k) - 1ull)); }
    inline void drop(int k) { buf >>= k; cnt -= k; }
    inline uint32_t take(int k) { const uint32_t v = peek(k); drop(k); return v; }
    uint64_t bitpos() const { return (uint64_t)pos * 8u - (uint64_t)cnt; }
    bool overrun() const { return bitpos() > (uint64_t)n * 8u; }
};

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline Entry symbol_entry(int kind, int sym, int bits)         // kind 0 literal / length alphabet, 1 distance alphabet
{
    Entry e; e.bits = (uint8_t)bits;
    if (kind == 0) {
        if (sym < 256) { e.val = (uint16_t)sym; e.op = 0; }
        else if (sym == 256) { e.val = 0; e.op = 32; }
        else if (sym < 286) { e.val = LEN_BASE[sym - 257]; e.op = (uint8_t)(16 + LEN_EXTRA[sym - 257]); }
        else { e.val = 0; e.op = 128; }
    } else {
        if (sym < 30) { e.val = DIST_BASE[sym]; e.op = (uint8_t)(16 + DIST_EXTRA[sym]); }
        else { e.val = 0; e.op = 128; }
    }
    return e;
}

// Canonical Huffman code (RFC 1951 3.2.2) -> two-level table indexed by the next bits of the stream.
// Returns -1 for an over-subscribed code, 0 for an incomplete one (unassigned patterns decode to "not a code"),
// 1 for a complete one.
inline int build_table(const uint8_t *lens, int n_sym, int kind, int prim, Entry *tab, int tab_cap)
{
    int count[16] = {0};
    for (int s = 0; s < n_sym; ++s) ++count[lens[s]];
    count[0] = 0;
    int left = 1;
    for (int l = 1; l <= 15; ++l) { left = 2 * left - count[l]; if (left < 0) return -1; }
    uint32_t next[16]; uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
    const Entry none = {0, 1, 128};
    for (int i = 0; i < (1 << prim); ++i) tab[i] = none;
    // first the short codes and, per primary index, the longest code that starts with it
    uint8_t sub_max[1 << LIT_PRIM];
    bool any_long = false;
    uint16_t rev_of[288];
    for (int s = 0; s < n_sym; ++s) {
        const int l = lens[s];
        if (!l) continue;
        uint32_t c = next[l]++, r = 0;
        for (int b = 0; b < l; ++b) { r = (r << 1) | (c & 1u); c >>= 1; }
        rev_of[s] = (uint16_t)r;
        if (l <= prim) {
            const Entry e = symbol_entry(kind, s, l);
            for (uint32_t i = r; i < (1u << prim); i += 1u << l) tab[i] = e;
        } else {
            if (!any_long) { memset(sub_max, 0, (size_t)1 << prim); any_long = true; }
            uint8_t &m = sub_max[r & ((1u << prim) - 1u)];
            if (l > m) m = (uint8_t)l;
        }
    }
    if (any_long) {
        int used = 1 << prim;
        for (int i = 0; i < (1 << prim); ++i) {
            if (!sub_max[i]) continue;
            const int sb = sub_max[i] - prim;
            if (used + (1 << sb) > tab_cap) return -1;
            Entry link; link.val = (uint16_t)used; link.bits = (uint8_t)prim; link.op = (uint8_t)(64 + sb);
            tab[i] = link;
            for (int k = 0; k < (1 << sb); ++k) tab[used + k] = none;
            used += 1 << sb;
        }
        for (int s = 0; s < n_sym; ++s) {
            const int l = lens[s];
            if (l <= prim) continue;
            const uint32_t r = rev_of[s];
            const Entry &link = tab[r & ((1u << prim) - 1u)];
            const int sb = link.op - 64;
            const Entry e = symbol_entry(kind, s, l - prim);
            for (uint32_t i = r >> prim; i < (1u << sb); i += 1u << (l - prim)) tab[link.val + i] = e;
        }
    }
    return left == 0 ? 1 : 0;
}

struct Tables {
    Entry lit[LIT_TABLE];
    Entry dist[DIST_TABLE];
};

// bytes a FASTQ / text file is made of: what a guessed block must decode to
inline bool text_byte(unsigned b) { return (b >= 32 && b < 127) || b == '\n' || b == '\r' || b == '\t'; }

enum { PG_OK = 0, PG_BAD = 1, PG_NOMEM = 2, PG_MORE = 3 };

// 16-bit symbol buffer with the unknown window in front: sym[0 .. WINDOW) are the markers, output starts at WINDOW
struct SymBuf {
    uint16_t *sym = nullptr; int64_t n = 0, cap = 0;         // n, cap: output symbols (the window is extra)
    ~SymBuf() { free(sym); }
    SymBuf() = default;
    SymBuf(const SymBuf &) = delete;
    SymBuf &operator=(const SymBuf &) = delete;
    bool reserve(int64_t want)
    {
        if (want <= cap) return true;
        int64_t c = cap ? cap : (1 << 20);
        while (c < want) c += c / 2 + (1 << 16);
        uint16_t *q = (uint16_t *)realloc(sym, sizeof(uint16_t) * (size_t)(WINDOW + c + 16));
        if (!q) return false;
#if defined(MADV_HUGEPAGE)
        static const bool thp = getenv("SCAR_PGZIP_THP") != nullptr;      // (experiment: transparent huge pages for the symbol buffers)
        if (thp) madvise((void *)(((uintptr_t)q + 4095) & ~(uintptr_t)4095), (sizeof(uint16_t) * (size_t)(WINDOW + c)) & ~(size_t)4095, MADV_HUGEPAGE);
#endif
        if (!sym) for (int i = 0; i < WINDOW; ++i) q[i] = (uint16_t)(256 + i);
        sym = q; cap = c;
        return true;
    }
    void release() { free(sym); sym = nullptr; n = cap = 0; }
    // a used allocation changes hands (the pages stay mapped: no first-touch faults for the next chunk)
    void adopt(uint16_t *p, int64_t c) { free(sym); sym = p; cap = c; n = 0; }
    uint16_t *give(int64_t &c) { uint16_t *p = sym; c = cap; sym = nullptr; n = cap = 0; return p; }
    uint16_t *out() { return sym + WINDOW; }
};

// the header of a dynamic block (RFC 1951 3.2.7) -> tables.  strict: both codes must be complete (a distance code of
// at most one symbol is allowed, as zlib writes it for a block with <= 1 distance), end-of-block must have a code.
inline bool read_dynamic_header(Bits &br, Tables &T)
{
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    br.refill();
    const int hlit = (int)br.take(5) + 257, hdist = (int)br.take(5) + 1, hclen = (int)br.take(4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    uint8_t cl[19] = {0};
    for (int i = 0; i < hclen; ++i) { if (br.cnt < 3) br.refill(); cl[order[i]] = (uint8_t)br.take(3); }
    Entry ct[128];
    {
        // code-length code: at most 7 bits, one level
        int count[8] = {0};
        for (int s = 0; s < 19; ++s) ++count[cl[s]];
        count[0] = 0;
        int left = 1;
        for (int l = 1; l <= 7; ++l) { left = 2 * left - count[l]; if (left < 0) return false; }
        if (left != 0) return false;
        if (build_table(cl, 19, 1, 7, ct, 128) < 0) return false;      // (kind 1 with sym < 30: val = DIST_BASE, unused; the symbol is recovered below)
    }
    // build_table stores bases; for the code-length alphabet the symbol itself is wanted: rebuild val = symbol
    {
        int count[8] = {0};
        for (int s = 0; s < 19; ++s) ++count[cl[s]];
        count[0] = 0;
        uint32_t next[8], code = 0;
        for (int l = 1; l <= 7; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
        for (int s = 0; s < 19; ++s) {
            const int l = cl[s];
            if (!l) continue;
            uint32_t c = next[l]++, r = 0;
            for (int b = 0; b < l; ++b) { r = (r << 1) | (c & 1u); c >>= 1; }
            for (uint32_t i = r; i < 128u; i += 1u << l) { ct[i].val = (uint16_t)s; ct[i].bits = (uint8_t)l; ct[i].op = 0; }
        }
    }
    uint8_t lens[286 + 30 + 138];
    int i = 0;
    const int total = hlit + hdist;
    while (i < total) {
        br.refill();
        const Entry e = ct[br.peek(7)];
        if (e.op) return false;
        br.drop(e.bits);
        const int s = e.val;
        if (s < 16) { lens[i++] = (uint8_t)s; continue; }
        int rep; uint8_t v = 0;
        if (s == 16) { if (i == 0) return false; v = lens[i - 1]; rep = 3 + (int)br.take(2); }
        else if (s == 17) rep = 3 + (int)br.take(3);
        else rep = 11 + (int)br.take(7);
        if (i + rep > total) return false;
        memset(lens + i, v, (size_t)rep);
        i += rep;
    }
    if (lens[256] == 0) return false;
    if (build_table(lens, hlit, 0, LIT_PRIM, T.lit, LIT_TABLE) != 1) return false;
    const int dc = build_table(lens + hlit, hdist, 1, DIST_PRIM, T.dist, DIST_TABLE);
    if (dc < 0) return false;
    if (dc == 0) {
        int used = 0;
        for (int s = 0; s < hdist; ++s) used += lens[hlit + s] != 0;
        if (used > 1) return false;
    }
    return !br.overrun();
}

inline const Tables &fixed_tables()
{
    static const Tables *T = [] {
        Tables *t = new Tables();
        uint8_t lens[288];
        for (int s = 0; s < 288; ++s) lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
        build_table(lens, 288, 0, LIT_PRIM, t->lit, LIT_TABLE);
        uint8_t dl[32];
        for (int s = 0; s < 32; ++s) dl[s] = 5;
        build_table(dl, 32, 1, DIST_PRIM, t->dist, DIST_TABLE);
        return t;
    }();
    return *T;
}

// The symbols of one Huffman block.  CHECK: a guessed block - literals must be text bytes, at most `max_out` symbols.
// Returns PG_OK at the end-of-block symbol.
template <bool CHECK>
inline int decode_huffman_block(Bits &br, const Tables &T, SymBuf &B, int64_t max_out)
{
    uint16_t *out = B.out();
    int64_t n = B.n, room = B.cap - 272;
    for (;;) {
        if (n > room || (CHECK && n > max_out)) {
            B.n = n;
            if (CHECK && n > max_out) return PG_MORE;
            if (br.overrun()) return PG_BAD;                // (a truncated stream decodes zeros for ever)
            if (!B.reserve(n + (1 << 20))) return PG_NOMEM;
            out = B.out(); room = B.cap - 272;
        }
        br.refill();
        Entry e = T.lit[br.peek(LIT_PRIM)];
        if (e.op == 0) {                                   // up to two literals off the primary table per refill
            if (CHECK && !text_byte(e.val)) return PG_BAD;
            out[n++] = e.val; br.drop(e.bits);
            e = T.lit[br.peek(LIT_PRIM)];
            if (e.op == 0) {
                if (CHECK && !text_byte(e.val)) return PG_BAD;
                out[n++] = e.val; br.drop(e.bits);
                e = T.lit[br.peek(LIT_PRIM)];
            }
            br.refill();
        }
        if (e.op & 64) { br.drop(LIT_PRIM); e = T.lit[e.val + br.peek(e.op - 64)]; }
        if (e.op == 0) {
            if (CHECK && !text_byte(e.val)) return PG_BAD;
            out[n++] = e.val; br.drop(e.bits);
            continue;
        }
        if (e.op & 128) return PG_BAD;
        br.drop(e.bits);
        if (e.op == 32) break;
        const int xb = e.op - 16;
        const int len = (int)e.val + (int)br.take(xb);
        Entry d = T.dist[br.peek(DIST_PRIM)];
        if (d.op & 64) { br.drop(DIST_PRIM); d = T.dist[d.val + br.peek(d.op - 64)]; }
        if (d.op & 128) return PG_BAD;
        br.drop(d.bits);
        const int64_t dist = (int64_t)d.val + (int64_t)br.take(d.op - 16);
        // (n - dist >= -WINDOW always: the markers stand in front of the output)
        uint16_t *dst = out + n;
        const uint16_t *src = dst - dist;
        n += len;
        if (dist >= 8) {
            for (int k = 0; k < len; k += 8) memcpy(dst + k, src + k, 16);
        } else if (dist >= 4) {
            for (int k = 0; k < len; k += 4) memcpy(dst + k, src + k, 8);
        } else if (dist == 1) {
            const uint64_t v = 0x0001000100010001ull * (uint64_t)src[0];
            for (int k = 0; k < len; k += 4) memcpy(dst + k, &v, 8);
        } else {
            for (int k = 0; k < len; ++k) dst[k] = src[k];
        }
    }
    B.n = n;
    if (br.overrun()) return PG_BAD;
    return PG_OK;
}

inline int decode_stored_block(Bits &br, SymBuf &B)
{
    br.drop(br.cnt & 7);
    br.refill();
    const uint32_t len = br.take(16), nlen = br.take(16);
    if ((len ^ nlen) != 0xFFFFu) return PG_BAD;
    // rewind to the byte position: everything still buffered is whole bytes
    int64_t at = (int64_t)(br.bitpos() >> 3);
    if (at + (int64_t)len > br.n) return PG_BAD;
    if (!B.reserve(B.n + (int64_t)len + 512)) return PG_NOMEM;
    uint16_t *out = B.out() + B.n;
    for (uint32_t k = 0; k < len; ++k) out[k] = br.p[at + k];
    B.n += len;
    br.init(br.p, br.n, (uint64_t)(at + (int64_t)len) * 8u);
    return PG_OK;
}

// One block at the reader's position.  Returns PG_OK and sets `final`.
inline int decode_block(Bits &br, Tables &T, SymBuf &B, bool &final)
{
    br.refill();
    final = br.take(1) != 0;
    const uint32_t type = br.take(2);
    if (type == 0) return decode_stored_block(br, B);
    if (type == 1) return decode_huffman_block<false>(br, fixed_tables(), B, 0);
    if (type == 3) return PG_BAD;
    if (!read_dynamic_header(br, T)) return PG_BAD;
    return decode_huffman_block<false>(br, T, B, 0);
}

// First bit position in [from, to) where a non-final dynamic block plausibly starts; UINT64_MAX when none does.
inline uint64_t find_block_start(const uint8_t *in, int64_t n, uint64_t from, uint64_t to, Tables &T, SymBuf &scratch)
{
    const uint64_t last = (uint64_t)n * 8u;
    if (to > last) to = last;
    for (uint64_t bit = from; bit + 64 < to + 64 && bit < to; ++bit) {
        // cheap rejects on the 17 header bits: BFINAL = 0, BTYPE = 10, HLIT <= 29, HDIST <= 29
        const int64_t byte = (int64_t)(bit >> 3);
        if (byte + 8 > n) break;
        uint64_t w;
        memcpy(&w, in + byte, 8);
        w >>= bit & 7;
        if ((w & 7u) != 4u) continue;                              // bits: final 0, type (LSB first) 0 1 -> value 2
        if (((w >> 3) & 31u) > 29u || ((w >> 8) & 31u) > 29u) continue;
        if (byte + 17 <= n) {
            // the code-length code must be complete: sum of 2^(7 - length) over its (HCLEN + 4) three-bit lengths = 128
            uint64_t hi;
            memcpy(&hi, in + byte + 8, 8);
            const int sh = (int)(bit & 7);
            // bits 17 .. 80 of the stream from `bit` (w holds its first 64 - sh bits)
            const uint64_t s_lo = sh ? (w | (hi << (64 - sh))) : w, s_hi = hi >> sh;
            uint64_t v = (s_lo >> 17) | (s_hi << 47);
            const int hclen = (int)((w >> 13) & 15u) + 4;
            int sum = 0;
            for (int i = 0; i < hclen; ++i) { const int l = (int)(v & 7u); v >>= 3; sum += (128 >> l) & 127; }
            if (sum != 128) continue;
        }
        Bits br;
        br.init(in, n, bit + 3);
        if (!read_dynamic_header(br, T)) continue;
        scratch.n = 0;
        if (!scratch.reserve(1 << 16)) return UINT64_MAX;
        // the first 32 K symbols of the block (or all of it, then a block header must follow)
        const int rc = decode_huffman_block<true>(br, T, scratch, 1 << 15);
        if (rc == PG_MORE) return bit;
        if (rc != PG_OK || scratch.n < 64) continue;               // a real block of a large file holds thousands of symbols
        br.refill();
        if (((br.peek(3) >> 1) & 3u) == 3u) continue;
        return bit;
    }
    return UINT64_MAX;
}

// gzip member header (RFC 1952) at byte `at`; returns the first byte of the DEFLATE data or -1
inline int64_t skip_gzip_header(const uint8_t *p, int64_t n, int64_t at)
{
    if (at + 18 > n || p[at] != 0x1f || p[at + 1] != 0x8b || p[at + 2] != 8) return -1;
    const int flg = p[at + 3];
    if (flg & 0xE0) return -1;
    int64_t q = at + 10;
    if (flg & 4) { if (q + 2 > n) return -1; q += 2 + (p[q] | (p[q + 1] << 8)); }
    if (flg & 8) { while (q < n && p[q]) ++q; ++q; }
    if (flg & 16) { while (q < n && p[q]) ++q; ++q; }
    if (flg & 2) q += 2;
    return q < n ? q : -1;
}

// markers -> bytes with the chunk's window; returns nothing, never fails (a marker always names a window position)
inline void translate(const uint16_t *sym, int64_t n, const uint8_t *window, uint8_t *dst)
{
    int64_t i = 0;
#if defined(__SSE2__)
    const __m128i hi = _mm_set1_epi16((short)0xFF00);
    for (; i + 16 <= n; i += 16) {
        const __m128i a = _mm_loadu_si128((const __m128i *)(sym + i)), b = _mm_loadu_si128((const __m128i *)(sym + i + 8));
        if (_mm_movemask_epi8(_mm_cmpeq_epi16(_mm_and_si128(_mm_or_si128(a, b), hi), _mm_setzero_si128())) == 0xFFFF) {
            _mm_storeu_si128((__m128i *)(dst + i), _mm_packus_epi16(a, b));
        } else {
            for (int k = 0; k < 16; ++k) { const uint16_t s = sym[i + k]; dst[i + k] = s < 256 ? (uint8_t)s : window[s - 256]; }
        }
    }
#endif
    for (int64_t k = 0, left = n - i; k < left; ++k) { const uint16_t s = sym[i + k]; dst[i + k] = s < 256 ? (uint8_t)s : window[s - 256]; }
}

struct MemberEnd { int64_t out_at; uint32_t crc, isize; };          // a member ended after out_at symbols of the chunk

struct Chunk {
    int64_t in_begin = 0, in_end = 0;          // nominal compressed byte range
    std::atomic<int> find_state{0};            // 0 not looked for, 1 being looked for, 2 known
    uint64_t start_bit = UINT64_MAX;           // guessed (chunk 0: true) block start, UINT64_MAX = none in range
    SymBuf out;
    uint8_t *window = nullptr;                 // the 32 KiB of text in front of this chunk (known in pass 2)
    uint64_t end_bit = 0;                      // the block boundary this chunk stopped on
    int64_t next_chunk = 0;                    // the chunk that starts there
    bool stream_end = false;                   // the input ended inside this chunk
    std::vector<MemberEnd> ends;
    int status = PG_OK; std::string error;
    std::atomic<int> done{0};
    ~Chunk() { free(window); }
};

class ParallelGunzip {
public:
    // in / n: the whole file (kept by the caller).  chunk_bytes <= 0: chosen from the file size and thread count.
    ParallelGunzip(const uint8_t *in, int64_t n, int n_threads, int64_t chunk_bytes) : in_(in), n_(n)
    {
        n_threads_ = std::max(1, n_threads);
        if (chunk_bytes <= 0) chunk_bytes = std::min<int64_t>(1 << 20, std::max<int64_t>(128 << 10, n / (8 * (int64_t)n_threads_)));
        chunk_bytes_ = chunk_bytes;
        data0_ = skip_gzip_header(in, n, 0);
        if (data0_ < 0) { fail_ = "not a gzip file (bad member header)"; return; }
        const int64_t n_chunks = std::max<int64_t>(1, (n - data0_ + chunk_bytes - 1) / chunk_bytes);
        chunks_ = std::vector<Chunk>((size_t)n_chunks);
        for (int64_t k = 0; k < n_chunks; ++k) {
            chunks_[(size_t)k].in_begin = data0_ + k * chunk_bytes;
            chunks_[(size_t)k].in_end = std::min(n, data0_ + (k + 1) * chunk_bytes);
        }
        chunks_[0].start_bit = expect_bit_ = (uint64_t)data0_ * 8u;
        chunks_[0].find_state = 2;
        lookahead_ = 2 * n_threads_ + 2;
        if (const char *e = getenv("SCAR_PGZIP_BUDGET")) { const long long v = atoll(e); if (v > 0) budget_ = v; }
        for (int t = 0; t < n_threads_; ++t) threads_.emplace_back([this] { worker(); });
    }
    ~ParallelGunzip()
    {
        { std::lock_guard<std::mutex> lk(m_); quit_ = true; }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
        for (auto &b : pool_) free(b.first);
    }
    const std::string &error() const { return fail_; }
    int threads() const { return n_threads_; }
    int64_t chunks_total() const { return (int64_t)chunks_.size(); }
    int64_t chunks_guessed() const { return guessed_ok_; }       // chunks whose guessed start was the true one
    int64_t chunks_redone() const { return redone_; }

    // Up to `cap` bytes of text into dst, stream order.  0 = end of the stream, -1 = error (see error()).
    // Returns early (with what there is) rather than wait for a chunk that is still being decoded.  The symbols of
    // everything handed out in one call are turned into bytes side by side (the calling thread takes part).
    int64_t read(uint8_t *dst, int64_t cap)
    {
        if (!fail_.empty()) return -1;
        int64_t got = 0;
        std::vector<Task> tasks;
        int64_t first_left = cur_;
        while (got < cap && !finished_) {
            Chunk &c = chunks_[(size_t)cur_];
            if (!cur_ready_) {
                if (got > 0 && !c.done.load(std::memory_order_acquire)) break;
                wait_done(c);
                if (got > 0 && needs_redo(c)) break;                               // (its buffer may still feed tasks of this call)
                if (!accept_current()) break;
            }
            const int64_t take = std::min(cap - got, c.out.n - cur_off_);
            for (int64_t o = 0; o < take;) {                                       // tasks of at most 512 K symbols, cut at member ends
                int64_t len = std::min<int64_t>(take - o, 1 << 19);
                while (end_idx_ < c.ends.size() && c.ends[end_idx_].out_at <= cur_off_ + o) close_member(c.ends[end_idx_++], tasks);
                if (end_idx_ < c.ends.size()) len = std::min(len, c.ends[end_idx_].out_at - (cur_off_ + o));
                Task t; t.sym = c.out.out() + cur_off_ + o; t.n = len; t.dst = dst + got + o; t.window = c.window; t.crc = 0; t.member_end = nullptr;
                tasks.push_back(t);
                o += len;
            }
            got += take; cur_off_ += take;
            if (cur_off_ == c.out.n) {
                while (end_idx_ < c.ends.size()) close_member(c.ends[end_idx_++], tasks);
                step_to_next();
            }
        }
        run_tasks(tasks);
        if (fail_.empty()) fold_crcs(tasks);
        release_until(first_left, cur_);
        return fail_.empty() ? got : -1;
    }

private:
    struct Task { const uint16_t *sym; int64_t n; uint8_t *dst; const uint8_t *window; uint32_t crc; const MemberEnd *member_end; };

    const uint8_t *in_; int64_t n_;
    int n_threads_ = 1; int64_t chunk_bytes_ = 0, data0_ = 0;
    std::vector<Chunk> chunks_;
    std::vector<std::thread> threads_;
    std::mutex m_; std::condition_variable cv_;
    bool quit_ = false;
    int64_t next_decode_ = 0, lookahead_ = 4;
    int64_t budget_ = 128ll << 20;                 // symbols a chunk may hold before it hands the stream back
    std::atomic<int64_t> consumed_{0};             // chunks the reader has left behind
    // translation tasks of the current read() (workers help)
    std::vector<Task> *tasks_ = nullptr; size_t task_next_ = 0, task_done_ = 0;
    std::atomic<int> busy_{0};                     // translation tasks are waiting for hands
    std::vector<std::pair<uint16_t *, int64_t>> pool_;   // symbol buffers of chunks left behind
    // reader state
    std::string fail_;
    int64_t cur_ = 0, cur_off_ = 0; bool cur_ready_ = false, finished_ = false;
    size_t end_idx_ = 0;
    uint8_t next_window_[WINDOW];
    uint64_t expect_bit_ = 0;
    uint32_t crc_run_ = 0; int64_t member_len_ = 0;
    int64_t guessed_ok_ = 0, redone_ = 0;
    const bool trace_ = getenv("SCAR_PGZIP_TRACE") != nullptr;
    const double t_begin_ = now_s();
    static double now_s()
    {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
    }

    // ---- pass 1 ----------------------------------------------------------------------------------------------------
    uint64_t start_of(int64_t k, Tables &T, SymBuf &scratch)            // the guessed start of chunk k (looked for once)
    {
        Chunk &c = chunks_[(size_t)k];
        int s = c.find_state.load(std::memory_order_acquire);
        if (s == 0) {
            int zero = 0;
            if (c.find_state.compare_exchange_strong(zero, 1)) {
                c.start_bit = find_block_start(in_, n_, (uint64_t)c.in_begin * 8u, (uint64_t)c.in_end * 8u, T, scratch);
                c.find_state.store(2, std::memory_order_release);
                return c.start_bit;
            }
        }
        while (c.find_state.load(std::memory_order_acquire) != 2) std::this_thread::yield();
        return c.start_bit;
    }

    // decode chunk k from `from` until the block boundary where a later chunk starts (or the end of the input)
    void decode_chunk(int64_t k, uint64_t from, Tables &T, SymBuf &scratch)
    {
        Chunk &c = chunks_[(size_t)k];
        c.ends.clear(); c.out.n = 0; c.status = PG_OK; c.stream_end = false; c.error.clear();
        if (!c.out.sym) take_pooled(c.out);
        if (!c.out.reserve(1 << 20)) { c.status = PG_NOMEM; c.error = "out of host memory"; return; }
        Bits br;
        br.init(in_, n_, from);
        int64_t j = k + 1;
        const int64_t nchunks = (int64_t)chunks_.size();
        for (;;) {
            bool final = false;
            const int rc = decode_block(br, T, c.out, final);
            if (rc != PG_OK) { c.status = rc; c.error = rc == PG_NOMEM ? "out of host memory" : "corrupt DEFLATE data"; return; }
            if (final) {
                // member trailer, then the next member's header or the end of the file (zero padding skipped)
                br.drop(br.cnt & 7);
                int64_t at = (int64_t)(br.bitpos() >> 3);
                if (at + 8 > n_) { c.status = PG_BAD; c.error = "unexpected end of file"; return; }
                MemberEnd me; me.out_at = c.out.n;
                me.crc = (uint32_t)in_[at] | ((uint32_t)in_[at + 1] << 8) | ((uint32_t)in_[at + 2] << 16) | ((uint32_t)in_[at + 3] << 24);
                me.isize = (uint32_t)in_[at + 4] | ((uint32_t)in_[at + 5] << 8) | ((uint32_t)in_[at + 6] << 16) | ((uint32_t)in_[at + 7] << 24);
                c.ends.push_back(me);
                at += 8;
                while (at < n_ && in_[at] == 0) ++at;
                if (at >= n_) { c.stream_end = true; c.end_bit = (uint64_t)n_ * 8u; c.next_chunk = nchunks; return; }
                const int64_t d0 = skip_gzip_header(in_, n_, at);
                if (d0 < 0) { c.status = PG_BAD; c.error = "bad gzip member header"; return; }
                br.init(in_, n_, (uint64_t)d0 * 8u);
            }
            const uint64_t p = br.bitpos();
            // has a later chunk started exactly here?
            bool stop = false;
            while (j < nchunks && p >= (uint64_t)chunks_[(size_t)j].in_begin * 8u) {
                const uint64_t s = start_of(j, T, scratch);
                if (s == p) { stop = true; break; }
                if (s != UINT64_MAX && s > p) break;                // its start lies ahead: keep decoding
                ++j;                                                // no start in its range, or one this stream jumped over
            }
            if (stop) { c.end_bit = p; c.next_chunk = j; return; }
            if (c.out.n > budget_ && k + 1 < nchunks) {
                // text that inflates beyond any FASTQ ratio: hand the stream back at this boundary so that memory stays
                // bounded.  The chunk that holds this bit (or the next one) is decoded again from here by the reader.
                const int64_t holder = ((int64_t)(p >> 3) - data0_) / chunk_bytes_;
                c.end_bit = p; c.next_chunk = std::min(std::max(holder, k + 1), nchunks - 1);
                return;
            }
            // between blocks: the reader's translation work comes first (it is what frees buffers and staging pieces)
            if (busy_.load(std::memory_order_acquire)) {
                std::unique_lock<std::mutex> lk(m_);
                if (quit_) { c.status = PG_BAD; c.error = "cancelled"; return; }
                help_translate(lk);
            }
        }
    }

    void take_pooled(SymBuf &b)
    {
        std::lock_guard<std::mutex> lk(m_);
        if (pool_.empty()) return;
        b.adopt(pool_.back().first, pool_.back().second);
        pool_.pop_back();
    }
    void give_pooled(SymBuf &b)
    {
        int64_t cap = 0;
        uint16_t *p = b.give(cap);
        if (!p) return;
        std::lock_guard<std::mutex> lk(m_);
        pool_.emplace_back(p, cap);
    }

    void worker()
    {
        Tables *T = new Tables();
        SymBuf scratch;
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            cv_.wait(lk, [&] {
                return quit_ || (tasks_ && task_next_ < tasks_->size()) ||
                       (next_decode_ < (int64_t)chunks_.size() && next_decode_ < consumed_.load() + lookahead_);
            });
            if (quit_) break;
            if (tasks_ && task_next_ < tasks_->size()) { help_translate(lk); continue; }
            const int64_t k = next_decode_++;
            lk.unlock();
            Chunk &c = chunks_[(size_t)k];
            const double t0 = trace_ ? now_s() : 0.0;
            const uint64_t s = start_of(k, *T, scratch);
            const double t1 = trace_ ? now_s() : 0.0;
            if (s != UINT64_MAX) decode_chunk(k, s, *T, scratch);
            else { c.status = PG_BAD; c.error = "no block start in range"; }
            if (trace_) fprintf(stderr, "[pgzip] chunk %3ld find %.1f ms decode %.1f ms (%.1f MB) started at %.1f ms status %d\n", (long)k, 1e3 * (t1 - t0), 1e3 * (now_s() - t1), c.out.n / 1e6, 1e3 * (t0 - t_begin_), c.status);
            c.done.store(1, std::memory_order_release);
            lk.lock();
            cv_.notify_all();
        }
        lk.unlock();
        delete T;
    }

    void help_translate(std::unique_lock<std::mutex> &lk)
    {
        while (tasks_ && task_next_ < tasks_->size()) {
            std::vector<Task> *ts = tasks_;
            Task &t = (*ts)[task_next_++];
            lk.unlock();
            translate(t.sym, t.n, t.window, t.dst);
            t.crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), t.dst, (uInt)t.n);
            lk.lock();
            if (++task_done_ == ts->size()) cv_.notify_all();
        }
    }

    void run_tasks(std::vector<Task> &tasks)
    {
        if (tasks.empty()) return;
        std::unique_lock<std::mutex> lk(m_);
        tasks_ = &tasks; task_next_ = 0; task_done_ = 0;
        busy_.store(1, std::memory_order_release);
        cv_.notify_all();
        help_translate(lk);
        busy_.store(0, std::memory_order_release);
        cv_.wait(lk, [&] { return task_done_ == tasks.size(); });
        tasks_ = nullptr;
    }

    // ---- pass 2 ----------------------------------------------------------------------------------------------------
    void wait_done(Chunk &c)
    {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return c.done.load(std::memory_order_acquire) != 0; });
    }

    // the current chunk has been decoded by a worker: take it if it started where the stream stands, else decode it
    // again from there (on this thread: rare)
    bool accept_current()
    {
        Chunk &c = chunks_[(size_t)cur_];
        if (needs_redo(c)) {
            ++redone_;
            Tables *T = new Tables();
            SymBuf scratch;
            c.start_bit = expect_bit_;
            decode_chunk(cur_, expect_bit_, *T, scratch);
            delete T;
        } else if (cur_ > 0) ++guessed_ok_;
        if (c.status != PG_OK) { fail_ = "gzip: " + (c.error.empty() ? std::string("corrupt stream") : c.error); return false; }
        c.window = (uint8_t *)malloc(WINDOW);
        if (!c.window) { fail_ = "gzip: out of host memory"; return false; }
        if (cur_ == 0) {
            // nothing stands in front of the first block: a marker here is a distance too far back
            const uint16_t *s = c.out.out();
            for (int64_t i = 0; i < c.out.n; ++i) if (s[i] >= 256) { fail_ = "gzip: invalid distance too far back"; return false; }
            memset(c.window, 0, WINDOW);
        } else memcpy(c.window, next_window_, WINDOW);
        // the window of whatever comes next: the last 32 KiB of the text up to the end of this chunk
        if (c.out.n >= WINDOW) translate(c.out.out() + c.out.n - WINDOW, WINDOW, c.window, next_window_);
        else {
            memcpy(next_window_, c.window + c.out.n, (size_t)(WINDOW - c.out.n));
            translate(c.out.out(), c.out.n, c.window, next_window_ + (WINDOW - c.out.n));
        }
        cur_off_ = 0; end_idx_ = 0; cur_ready_ = true;
        return true;
    }

    bool needs_redo(const Chunk &c) const { return c.start_bit != expect_bit_ || c.status == PG_BAD; }

    void step_to_next()
    {
        Chunk &c = chunks_[(size_t)cur_];
        expect_bit_ = c.end_bit;
        cur_ready_ = false;
        if (c.stream_end || c.next_chunk >= (int64_t)chunks_.size()) {
            finished_ = true;
            if (!c.stream_end) fail_ = "gzip: unexpected end of file";
        }
        cur_ = c.next_chunk;
    }

    // chunks [from, to) are behind the reader and their symbols have become bytes: buffers back to the pool.  A chunk
    // in between that was decoded on a wrong guess is waited for (or never started) and dropped.
    void release_until(int64_t from, int64_t to)
    {
        to = std::min<int64_t>(to, (int64_t)chunks_.size());
        if (to <= from) return;
        int64_t started;
        {
            std::lock_guard<std::mutex> lk(m_);
            started = next_decode_;
            if (next_decode_ < to) next_decode_ = to;
        }
        for (int64_t k = from; k < to; ++k) {
            Chunk &c = chunks_[(size_t)k];
            if (k < started) wait_done(c);
            give_pooled(c.out);
            free(c.window); c.window = nullptr;
        }
        { std::lock_guard<std::mutex> lk(m_); consumed_.store(to); }
        cv_.notify_all();
    }

    void close_member(const MemberEnd &me, std::vector<Task> &tasks)
    {
        // marks the boundary inside the task list: the CRC fold checks the member when it passes this point
        Task t; t.sym = nullptr; t.n = 0; t.dst = nullptr; t.window = nullptr; t.crc = 0; t.member_end = &me;
        tasks.push_back(t);
    }

    bool check_member(const MemberEnd &me)
    {
        if (crc_run_ != me.crc || (uint32_t)member_len_ != me.isize) { fail_ = "gzip: CRC or length of a member does not match its trailer"; return false; }
        crc_run_ = 0; member_len_ = 0;
        return true;
    }

    bool fold_crcs(const std::vector<Task> &tasks)
    {
        for (const Task &t : tasks) {
            if (t.member_end) { if (!check_member(*t.member_end)) return false; continue; }
            crc_run_ = member_len_ ? (uint32_t)crc32_combine(crc_run_, t.crc, (z_off_t)t.n) : t.crc;
            member_len_ += t.n;
        }
        return true;
    }
};

}   // namespace scar_pgzip
