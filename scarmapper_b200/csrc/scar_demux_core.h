// scar_demux_core.h — index-matching arithmetic shared by the K2 kernel and the host unit tests
// (compiled for the host by tests/test_host_logic.py to check the bit-vector recurrence against
// a plain DP without a GPU).
#pragma once
#include <stdint.h>
#include "../../include/scar_b200.h"

#ifdef __CUDACC__
#define SCAR_HD __host__ __device__ __forceinline__
#else
#define SCAR_HD static inline
#endif
#if defined(__CUDA_ARCH__)
#define SCAR_POPCLL(x) __popcll(x)
#else
#define SCAR_POPCLL(x) __builtin_popcountll(x)
#endif

struct DemuxString {           // one distinct index string
    uint64_t code[2];          // 4-bit class codes, symbol k at bits 4k (k < 32)
    int32_t len;
    int32_t peq_off;           // into peq[]: (n_classes) 64-bit masks, class c at peq_off + c
};


// Global Levenshtein distance between the pattern (Peq masks, length m <= 64) and the text.
SCAR_HD int myers_global(const uint64_t *peq, int m, const uint8_t *text, int n,
                                            const uint8_t *class_map)
{
    if (m == 0) return n;
    uint64_t Pv = ~0ull, Mv = 0ull;
    const uint64_t top = 1ull << (m - 1);
    int score = m;
    for (int j = 0; j < n; ++j) {
        const uint64_t Eq = peq[class_map[text[j]]];
        const uint64_t Xv = Eq | Mv;
        const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
        uint64_t Ph = Mv | ~(Xh | Pv);
        uint64_t Mh = Pv & Xh;
        score += (Ph & top) ? 1 : 0;
        score -= (Mh & top) ? 1 : 0;
        Ph = (Ph << 1) | 1ull;           // global distance: D[0][j] = j
        Mh <<= 1;
        Pv = Mh | ~(Xv | Ph);
        Mv = Ph & Xv;
    }
    return score;
}

// item r of a ragged list: CSR (offsets only), views (offsets + lengths: item r starts at offsets[r] and is
// lengths[r] long, e.g. fields inside a FASTQ text buffer), or fixed stride
SCAR_HD void get_item(const scar_ragged &R, int64_t r, const uint8_t *&p, int &len)
{
    if (R.offsets) {
        const int64_t s = R.offsets[r];
        p = R.bytes + s;
        len = R.lengths ? R.lengths[r] : (int)(R.offsets[r + 1] - s);
    } else { p = R.bytes + r * R.stride; len = R.lengths ? R.lengths[r] : (int)R.stride; }
}

SCAR_HD void class_codes(const uint8_t *u, int n, const uint8_t *class_map, uint64_t code[2])
{
    code[0] = code[1] = 0ull;
    for (int k = 0; k < n && k < 32; ++k)
        code[k >> 4] |= (uint64_t)class_map[u[k]] << (4 * (k & 15));
}

SCAR_HD int hamming_classes(const uint64_t a[2], const uint64_t b[2])
{
    int h = 0;
    for (int i = 0; i < 2; ++i) {
        uint64_t x = a[i] ^ b[i];
        x = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x1111111111111111ull;
        h += SCAR_POPCLL(x);
    }
    return h;
}

// does match_maker(d, u) <= t hold?
// use_codes = false skips the class-code Hamming shortcut (ucode not available / already decided)
SCAR_HD bool accepts(const DemuxString &d, const uint64_t *peq, const uint8_t *u, int n,
                                        const uint64_t ucode[2], const uint8_t *class_map, int t,
                                        bool use_codes = true)
{
    const int m = d.len;
    if (use_codes && n == m && m <= 32) {
        // class 0 ("other") never equals an index symbol (classes >= 1) so the XOR test is exact
        const int h = hamming_classes(d.code, ucode);
        if (h <= t) return true;          // lev <= hamming
        if (t <= 1) return false;         // equal lengths: lev <= 1 iff hamming <= 1
    }
    if (n < m && 2 * (m - n) > t) return false;   // lev >= m-n, so lev + (m-n) > t
    const int lev = myers_global(peq + d.peq_off, m, u, n, class_map);
    return lev - (n - m) <= t;
}

