// scar_swar.h — SWAR byte classification shared by K1 (pack) and K2 (demux): four read bytes per 32-bit
// word -> 2-bit base codes and ACGT / 'N' masks.  Host + device, so that tests/test_host_logic.py can check
// every formula exhaustively against the byte-at-a-time definition without a GPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SCAR_SWAR __host__ __device__ __forceinline__
#else
#define SCAR_SWAR static inline
#endif

// 0x80 in every byte of x that is zero (exact, no borrow propagation)
SCAR_SWAR uint32_t zero_bytes(uint32_t x)
{
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
}

// canonical-letter difference of four bytes.  The code (w>>1)&3 names the only letter a byte could be; t marks the
// bytes whose code is 2 (bits 2,1 = 1,0: 'T' if anything).  Outside bits 1 and 2 an A, C or G is 0x41 and a T is
// 0x50 = 0x41 + 15, so  raw = w ^ (0x41414141 + 15 t)  has (byte k & 0xF9) == 0 iff byte k of w is A, C, G or T
// ('E' = 0x45 has t = 1 and fails on 0x50; 'P','R','V' = 0x50,0x52,0x56 have t = 0 and fail on 0x41).
SCAR_SWAR uint32_t acgt_raw4(uint32_t w)
{
    const uint32_t t = ((w >> 2) & ~(w >> 1)) & 0x01010101u;            // 1 where the code is 2
    return w ^ (t * 15u + 0x41414141u);
}

// byte k of the result is 0 iff byte k of w is A, C, G or T
SCAR_SWAR uint32_t acgt_diff4(uint32_t w) { return acgt_raw4(w) & 0xF9F9F9F9u; }

// non-zero iff any of the sixteen bytes is not A, C, G or T (one mask for the four words)
SCAR_SWAR uint32_t acgt_any16(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3)
{
#ifdef SCAR_OLD_ACGT_TEST          // (A/B: the per-word test of round 2's first fused kernel)
    return acgt_diff4(v0) | acgt_diff4(v1) | acgt_diff4(v2) | acgt_diff4(v3);
#else
    return (acgt_raw4(v0) | acgt_raw4(v1) | acgt_raw4(v2) | acgt_raw4(v3)) & 0xF9F9F9F9u;
#endif
}

// 0x80 flags of four bytes -> 4-bit mask (one multiply: the four partial products land on distinct bits)
SCAR_SWAR uint32_t flags_to_nibble(uint32_t v) { return (v * 0x00204081u) >> 28; }

// 2-bit codes of four bytes (A=0 C=1 T=2 G=3 = (b>>1)&3) gathered into the TOP byte by one multiply
SCAR_SWAR uint32_t codes_to_top_byte(uint32_t w) { return ((w >> 1) & 0x03030303u) * 0x01041040u; }

// four bytes -> 8-bit codes, 4-bit ACGT mask
SCAR_SWAR void classify4(uint32_t w, uint32_t &code8, uint32_t &valid4)
{
    code8 = codes_to_top_byte(w) >> 24;
    const uint32_t x = acgt_diff4(w);
    valid4 = x ? flags_to_nibble(zero_bytes(x)) : 0xFu;
}
