// scar_pack.cu — K1: the packed read loader.
//
// Raw read bytes (as the FASTQ parser delivers them: FASTQ_Reader.seq_read, reference
// Valkyries/FASTQ_Tools.py:623-653) -> fixed-stride rows of 64-bit cells in HBM: 2-bit base codes,
// an ACGT mask and an 'N' mask (scar_internal.cuh).  The platform's stored-sequence transform is
// applied on the way (INDEL_Processing.py:946 Illumina as is; :978,1181-1182 TruSeq seq[6:-6];
// :981 Ramsden rcomp(seq)), so K3 always sees the sequence ScarSearch would be handed.
//
// One thread packs one read, one cell (16 bases) at a time: up to five aligned 32-bit loads covering
// the cell's 16-byte window, a funnel-shift realignment, then SWAR classification of four bytes at a
// time.  The output is cell-major ([cell][read]) so that each store of a warp is one 256-byte line and
// the one-thread-per-read kernels downstream load coalesced.
#include "scar_internal.cuh"

__global__ void __launch_bounds__(256) scar_pack_kernel(const K1Params P)
{
    // a CTA owns a tile of blockDim.x consecutive reads and walks their cells j = 0..W-1: the tile's
    // raw bytes (~38 KB at 150 nt) stay in L1 across the j loop, the stores of one j are coalesced
    const int64_t tiles = (P.n_reads + blockDim.x - 1) / blockDim.x;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t r = tile * blockDim.x + threadIdx.x;
        if (r >= P.n_reads) continue;
        const uint8_t *rp; int len;
        get_item(P.seq, r, rp, len);
        const int64_t start = rp - P.seq.bytes;
        int slen = len, shift = 0;
        if (P.platform == SCAR_PLATFORM_TRUSEQ) { slen = len > 12 ? len - 12 : 0; shift = 6; }
        if (slen > 16 * P.W) { atomicOr(P.status, 1); slen = 16 * P.W; }
        const bool rc = P.platform == SCAR_PLATFORM_RAMSDEN;
        const uintptr_t base = (uintptr_t)P.seq.bytes;
        int n_count = 0;
        for (int j = 0; j < P.W; ++j) {
            uint64_t *out = P.cells + (int64_t)j * P.stride + r;
            int nb = slen - 16 * j;                    // bases of this cell
            if (nb <= 0) { *out = 0ull; continue; }
            if (nb > 16) nb = 16;
            // 16-byte window of raw bytes: forward [ws, ws+16) holds stored bases 16j.. in order;
            // reverse-complement: window ends at raw index len-16j, stored base k = comp(window[15-k])
            const int64_t ws = rc ? start + (len - 16 * j) - 16 : start + shift + 16 * j;
            const int64_t lo = rc ? start + (len - 16 * j) - nb : ws;          // first needed byte
            const int64_t hi = rc ? start + (len - 16 * j) : ws + nb;          // one past the last
            const uintptr_t a = base + (uintptr_t)ws;
            const uintptr_t a0 = a & ~(uintptr_t)3;
            const uint32_t sh = (uint32_t)(a & 3) * 8;
            uint32_t w[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const uintptr_t wa = a0 + 4 * i;
                // load only words that contain a needed byte (never touches memory outside the read)
                const bool need = wa + 4 > base + (uintptr_t)lo && wa < base + (uintptr_t)hi;
                w[i] = need ? __ldg(reinterpret_cast<const uint32_t *>(wa)) : 0u;
            }
            uint32_t v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], sh);
            if (rc) {   // reverse the 16 bytes
                const uint32_t r0 = __byte_perm(v[3], 0, 0x0123), r1 = __byte_perm(v[2], 0, 0x0123);
                const uint32_t r2 = __byte_perm(v[1], 0, 0x0123), r3 = __byte_perm(v[0], 0, 0x0123);
                v[0] = r0; v[1] = r1; v[2] = r2; v[3] = r3;
            }
            uint32_t code = 0, valid = 0, nmask = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t c8, v4, n4;
                classify4(v[i], c8, v4, n4);
                code |= c8 << (8 * i); valid |= v4 << (4 * i); nmask |= n4 << (4 * i);
            }
            if (rc) code ^= 0xAAAAAAAAu;                               // complement: code ^ 2
            const uint32_t bm = nb == 16 ? 0xFFFFu : ((1u << nb) - 1u);
            valid &= bm; nmask &= bm;
            n_count += __popc(nmask);
            // spread the ACGT mask to 2 bits per base to clear the codes of non-ACGT / absent bases
            uint32_t m2 = valid;
            m2 = (m2 | (m2 << 8)) & 0x00FF00FFu; m2 = (m2 | (m2 << 4)) & 0x0F0F0F0Fu;
            m2 = (m2 | (m2 << 2)) & 0x33333333u; m2 = (m2 | (m2 << 1)) & 0x55555555u;
            code &= m2 * 3u;
            *out = (uint64_t)code | ((uint64_t)valid << 32) | ((uint64_t)nmask << 48);
        }
        P.lens[r] = (uint32_t)slen | ((uint32_t)n_count << 16);     // length and literal-'N' count for the filter
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
    const int64_t cap = (int64_t)sm_count * 8;          // 8 resident CTAs per SM, grid-stride over tiles
    if (blocks > cap) blocks = cap;
    scar_pack_kernel<<<(unsigned)blocks, threads, 0, stream>>>(P);
    if (launches) ++*launches;
    return cudaGetLastError();
}
