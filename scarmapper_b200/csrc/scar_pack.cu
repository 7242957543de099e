// scar_pack.cu — K1: the packed read loader.
//
// Raw read bytes (as the FASTQ parser delivers them: FASTQ_Reader.seq_read, reference
// Valkyries/FASTQ_Tools.py:623-653) -> fixed-stride rows of 64-bit cells in HBM: 2-bit base codes,
// an ACGT mask and an 'N' mask (scar_internal.cuh).  The platform's stored-sequence transform is
// applied on the way (INDEL_Processing.py:946 Illumina as is; :978,1181-1182 TruSeq seq[6:-6];
// :981 Ramsden rcomp(seq)), so K3 always sees the sequence ScarSearch would be handed.
//
// A CTA takes tiles of 256 consecutive reads; the tile's raw bytes (one contiguous span of the input)
// are brought into shared memory by ONE TMA bulk copy, then one thread packs one read, one cell
// (16 bases) at a time: five aligned 32-bit loads covering the cell's 16-byte window, a funnel-shift
// realignment, then SWAR classification of four bytes at a time.  The output is cell-major
// ([cell][read]) so that each store of a warp is one 256-byte line and the one-thread-per-read
// kernels downstream load coalesced.
#include "scar_internal.cuh"

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define PACK_TILE_BYTES (55 * 1024)     // staged raw bytes of one 256-read tile; 4 CTAs per SM

// pack one read.  STAGED: the tile's raw bytes are in shared memory (word k of the read at sword + 4k),
// otherwise they are read from global memory through rp4.
template <bool STAGED>
__device__ __forceinline__ void pack_read(const K1Params &P, int64_t r, const uint32_t *rp4, uint32_t sword, int mis,
                                          int len, bool rc)
{
    int slen = len, shift = 0;
    if (P.platform == SCAR_PLATFORM_TRUSEQ) { slen = len > 12 ? len - 12 : 0; shift = 6; }
    if (slen > 16 * P.W) { atomicOr(P.status, 1); slen = 16 * P.W; }
    const int kend = (mis + len + 3) >> 2;          // aligned words that overlap the raw read
    int n_count = 0;
    uint64_t *out = P.cells + r;
    for (int j = 0; j < P.W; ++j, out += P.stride) {
        const int nb = slen - 16 * j;               // bases of this cell (may exceed 16)
        if (nb <= 0) { *out = 0ull; continue; }
        // 16-byte window of raw bytes: forward, window byte k is stored base 16j+k; reverse
        // complement, the window ends at raw index len-16j and stored base k = comp(window[15-k])
        const int q = mis + (rc ? len - 16 * j - 16 : shift + 16 * j);
        const int k0 = q >> 2;                      // floor, q < 0 only in the last reverse cell
        const uint32_t sh = (uint32_t)(q & 3) * 8;
        uint32_t w[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int k = k0 + i;
            w[i] = 0u;
            if (k >= 0 && k < kend) {               // never touches a word outside the read
                if (STAGED) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[i]) : "r"(sword + 4u * (uint32_t)k));
                else w[i] = __ldg(rp4 + k);
            }
        }
        uint32_t v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], sh);
        if (rc) {   // reverse the 16 bytes
            const uint32_t r0 = __byte_perm(v[3], 0, 0x0123), r1 = __byte_perm(v[2], 0, 0x0123);
            const uint32_t r2 = __byte_perm(v[1], 0, 0x0123), r3 = __byte_perm(v[0], 0, 0x0123);
            v[0] = r0; v[1] = r1; v[2] = r2; v[3] = r3;
        }
        // 2-bit codes: (w>>1)&3 per byte, gathered into the top byte by one multiply per word
        uint32_t p[4], x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            p[i] = codes_to_top_byte(v[i]);
            x[i] = acgt_diff4(v[i]);
        }
        uint32_t code = __byte_perm(__byte_perm(p[0], p[1], 0x0073), __byte_perm(p[2], p[3], 0x0073), 0x5410);
        if (rc) code ^= 0xAAAAAAAAu;                               // complement: code ^ 2
        uint32_t valid = 0xFFFFu, nmask = 0u;
        if (nb < 16 || (x[0] | x[1] | x[2] | x[3]) != 0u) {        // rare: last cell, or a non-ACGT byte
            valid = 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                valid |= flags_to_nibble(zero_bytes(x[i])) << (4 * i);
                nmask |= flags_to_nibble(zero_bytes(v[i] ^ 0x4E4E4E4Eu)) << (4 * i);
            }
            const uint32_t bm = nb >= 16 ? 0xFFFFu : ((1u << nb) - 1u);
            valid &= bm; nmask &= bm;
            n_count += __popc(nmask);
            // spread the ACGT mask to 2 bits per base to clear the codes of non-ACGT / absent bases
            uint32_t m2 = valid;
            m2 = (m2 | (m2 << 8)) & 0x00FF00FFu; m2 = (m2 | (m2 << 4)) & 0x0F0F0F0Fu;
            m2 = (m2 | (m2 << 2)) & 0x33333333u; m2 = (m2 | (m2 << 1)) & 0x55555555u;
            code &= m2 * 3u;
        }
        *out = (uint64_t)code | ((uint64_t)valid << 32) | ((uint64_t)nmask << 48);
    }
    P.lens[r] = (uint32_t)slen | ((uint32_t)n_count << 16);     // length and literal-'N' count for the filter
}

// smallest start / largest end address over the threads with `mine` (CTA-wide; contains two barriers)
__device__ __forceinline__ void span_reduce(bool mine, const uint8_t *rp, int len, unsigned long long (*s_red)[8],
                                            unsigned long long &lo, unsigned long long &hi)
{
    unsigned long long a = mine ? (unsigned long long)(uintptr_t)rp : ~0ull;
    unsigned long long b = mine ? (unsigned long long)(uintptr_t)rp + (unsigned long long)len : 0ull;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const unsigned long long a2 = __shfl_xor_sync(0xFFFFFFFFu, a, d), b2 = __shfl_xor_sync(0xFFFFFFFFu, b, d);
        a = a2 < a ? a2 : a; b = b2 > b ? b2 : b;
    }
    __syncthreads();                                  // s_red may still be read from the previous use
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = a; s_red[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    lo = s_red[0][0]; hi = s_red[1][0];
#pragma unroll
    for (int i = 1; i < 8; ++i) { lo = s_red[0][i] < lo ? s_red[0][i] : lo; hi = s_red[1][i] > hi ? s_red[1][i] : hi; }
}

__global__ void __launch_bounds__(256, 4) scar_pack_kernel(const K1Params P)
{
    // A CTA owns tiles of 256 consecutive reads.  The raw bytes of a tile are one contiguous span of the
    // input (CSR / fixed stride; also FASTQ views in file order): ONE TMA bulk copy brings the span into
    // shared memory (coalesced, no register traffic), then each thread packs its read out of shared
    // memory with word loads.  A span larger than the buffer (FASTQ text: the reads are a third of the
    // bytes) is taken in 2, 4 or 8 passes over 128, 64 or 32 reads; what still does not fit (very long
    // reads, scattered views) is read straight from global memory.  The stores of one cell index j are
    // coalesced either way.
    extern __shared__ __align__(128) unsigned char tile_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(tile_raw);
    unsigned char *tile = tile_raw + 16;
    __shared__ unsigned long long s_red[2][8];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    const bool rc = P.platform == SCAR_PLATFORM_RAMSDEN;
    const bool views = P.seq.offsets && P.seq.lengths;
    const int64_t tiles = (P.n_reads + blockDim.x - 1) / blockDim.x;
    for (int64_t tile_no = blockIdx.x; tile_no < tiles; tile_no += gridDim.x) {
        const int64_t r = tile_no * blockDim.x + threadIdx.x;
        const bool active = r < P.n_reads;
        const uint8_t *rp = P.seq.bytes; int len = 0;
        if (active) get_item(P.seq, r, rp, len);
        // span of reads [a, b] of the batch: monotonic layouts (CSR, fixed stride) need only the two end items
        // (thread 0 looks them up); views may be in any order and take a CTA-wide min / max
        const int64_t r0 = tile_no * blockDim.x;
        auto span_of = [&](int64_t a, int64_t b, bool mine, unsigned long long &lo, unsigned long long &hi) {
            if (views) { span_reduce(mine, rp, len, s_red, lo, hi); return; }
            if (threadIdx.x == 0) {
                if (b >= P.n_reads) b = P.n_reads - 1;
                unsigned long long x = ~0ull, y = 0ull;
                if (a <= b) {
                    const uint8_t *q0, *q1; int l0, l1;
                    get_item(P.seq, a, q0, l0);
                    get_item(P.seq, b, q1, l1);
                    x = (unsigned long long)(uintptr_t)q0; y = (unsigned long long)(uintptr_t)q1 + (unsigned long long)l1;
                }
                s_red[0][0] = x; s_red[1][0] = y;
            }
            __syncthreads();                          // (the barrier that ends a pass protects the reuse)
            lo = s_red[0][0]; hi = s_red[1][0];
        };
        unsigned long long lo, hi;
        span_of(r0, r0 + blockDim.x - 1, active, lo, hi);
        // passes: 1 when the whole tile fits, else the power of two that brings the average pass under the budget
        int passes = 1;
        {
            const unsigned long long whole = hi > lo ? hi - (lo & ~15ull) + 15ull : 0ull;
            while (passes < 8 && whole > (unsigned long long)passes * (PACK_TILE_BYTES - 64)) passes *= 2;
        }
        const int sub_shift = 8 - (31 - __clz(passes)), sub = 1 << sub_shift;      // blockDim.x = 256
        for (int ps = 0; ps < passes; ++ps) {
            if (passes > 1) __syncthreads();          // everyone has read the previous span
            const bool mine = active && ((int)threadIdx.x >> sub_shift) == ps;
            if (passes > 1) span_of(r0 + (int64_t)ps * sub, r0 + (int64_t)(ps + 1) * sub - 1, mine, lo, hi);
            const unsigned long long lo16 = lo & ~15ull;
            const unsigned long long span = hi > lo16 ? ((hi - lo16 + 15ull) & ~15ull) : 0ull;
            const bool staged = span != 0ull && span <= (unsigned long long)PACK_TILE_BYTES;
            if (staged) {
                if (threadIdx.x == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
                                 "r"((uint32_t)span) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_addr(tile)), "l"(lo16), "r"((uint32_t)span), "r"(smem_addr(bar)) : "memory");
                }
                uint32_t done = 0;
                while (!done) {
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                                 "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
                }
                parity ^= 1u;
            }
            if (mine) {
                const uintptr_t a4 = (uintptr_t)rp & ~(uintptr_t)3;
                const int mis = (int)((uintptr_t)rp & 3);
                if (staged) pack_read<true>(P, r, nullptr, smem_addr(tile) + (uint32_t)(a4 - lo16), mis, len, rc);
                else pack_read<false>(P, r, reinterpret_cast<const uint32_t *>(a4), 0u, mis, len, rc);
            }
            __syncthreads();                          // the tile buffer is reused by the next pass / tile
        }
    }
}

cudaError_t scar_launch_pack(const scar_ragged &seq, int64_t n, int W, int platform, uint64_t *cells,
                             int64_t stride, uint32_t *lens, int32_t *status, int sm_count, cudaStream_t stream,
                             int *launches)
{
    if (n <= 0) return cudaSuccess;
    K1Params P; P.seq = seq; P.cells = cells; P.stride = stride; P.lens = lens; P.n_reads = n; P.W = W;
    P.platform = platform; P.status = status;
    const int threads = 256;
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count * 4;          // 4 resident CTAs per SM, grid-stride over tiles
    if (blocks > cap) blocks = cap;
    const int smem = PACK_TILE_BYTES + 16;
    cudaError_t e = cudaFuncSetAttribute(scar_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;     // per device, cheap; a process may hold contexts on several devices
    scar_pack_kernel<<<(unsigned)blocks, threads, smem, stream>>>(P);
    if (launches) ++*launches;
    return cudaGetLastError();
}
