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

// canonical-letter difference of four bytes: byte k of the result is 0 iff byte k of w is A, C, G or T.
// The code (w>>1)&3 names the only letter the byte could be; 'A'|(w&6) is that letter for A, C, G and
// 'E' for code 2, which +15 turns into 'T'.
SCAR_SWAR uint32_t acgt_diff4(uint32_t w)
{
    const uint32_t t = ((w >> 2) & ~(w >> 1)) & 0x01010101u;            // 1 where the code is 2
    const uint32_t canon = (0x41414141u | (w & 0x06060606u)) + t * 15u;
    return w ^ canon;
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
