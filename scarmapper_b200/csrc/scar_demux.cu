// scar_demux.cu — K2: sample assignment by index edit distance.
//
// Replaces DataProcessing.index_matching (reference scarmapper/INDEL_Processing.py:1126-1201) and
// Sequence_Magic.match_maker (Valkyries/Sequence_Magic.py:123-136) for a whole batch:
//     match_maker(q, u) = Levenshtein(q, u) - (len(u) - len(q))
//     accept the FIRST manifest row whose two index sequences both score <= mismatch.
// The reference recomputes both distances for every manifest row; here the distinct index strings
// of each side are scored once per read (Illumina: 12 + 8 instead of 96 x 2), a bit mask of
// acceptable strings per side is formed, and the first manifest row is read from a
// [distinct-right x distinct-left] matrix.  The distance itself is the Myers/Hyyro bit-vector
// recurrence for the GLOBAL unit-cost edit distance (one 64-bit column per text character), with an
// exact shortcut for equal lengths: Hamming <= t implies accepted; for t <= 1 Hamming > t implies
// rejected (one edit that keeps the length is a substitution).
//
// Unknown strings per platform (u0 is scored against right_index = index_dict[s][2], u1 against
// left_index = index_dict[s][0]):
//   Illumina: u0, u1 = header fields f0, f1           (:1155-1156)
//   TruSeq:   u0 = seq[:6],  u1 = seq[-6:]             (:1160-1161)
//   Ramsden:  u0 = seq[:len(right_index)], u1 = seq[-len(left_index):]   (:1165,1170-1171)
#include "scar_internal.cuh"


// 2-bit codes + ACGT mask of a string of <= 16 bytes (aligned word loads + SWAR classification)
__device__ __forceinline__ void encode16(const uint8_t *u, int n, uint32_t &code, uint32_t &valid)
{
    const uintptr_t a = (uintptr_t)u, a0 = a & ~(uintptr_t)3;
    if (n == 8 && (a & 7) == 0) {                 // the usual Illumina field: one aligned 8-byte load
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(a));
        uint32_t c8, v4, c8b, v4b;
        classify4(v.x, c8, v4);
        classify4(v.y, c8b, v4b);
        code = c8 | (c8b << 8); valid = v4 | (v4b << 4);
        return;
    }
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t w[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const uintptr_t wa = a0 + 4 * i;
        w[i] = (wa < a + (uintptr_t)n) ? __ldg(reinterpret_cast<const uint32_t *>(wa)) : 0u;
    }
    code = 0; valid = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (4 * i < n) {                          // (8-nt Illumina fields: two words)
            uint32_t c8, v4;
            classify4(__funnelshift_r(w[i], w[i + 1], sh), c8, v4);
            code |= c8 << (8 * i); valid |= v4 << (4 * i);
        }
    }
    valid &= n >= 16 ? 0xFFFFu : ((1u << n) - 1u);
    code &= n >= 16 ? 0xFFFFFFFFu : ((1u << (2 * n)) - 1u);    // bytes past the string belong to the next item
}

// every position that is invalid or past the length counts as a mismatch: bit 2k set
__device__ __forceinline__ uint32_t spread_invalid(uint32_t valid, int n)
{
    uint32_t m = ~valid & (n >= 16 ? 0xFFFFu : ((1u << n) - 1u));
    if (m == 0u) return 0u;                       // the usual case: every base is A, C, G or T
    m = (m | (m << 8)) & 0x00FF00FFu; m = (m | (m << 4)) & 0x0F0F0F0Fu;
    m = (m | (m << 2)) & 0x33333333u; m = (m | (m << 1)) & 0x55555555u;
    return m;
}

template <bool FAST>
__global__ void __launch_bounds__(256) scar_demux_kernel(const K2Params P)
{
    __shared__ uint8_t cmap[256];
    __shared__ uint32_t s_code[2][64];     // 2-bit codes of the distinct index strings (FAST)
    __shared__ int32_t s_len[2][64];
    __shared__ unsigned long long s_lenmask[2][17];   // bit d set: string d of that side is n nt long (FAST)
    __shared__ uint16_t s_first[1024];                // first_sample matrix when it is small enough
    const int n_pairs = P.n_right * P.n_left;
    const bool first_in_smem = n_pairs <= 1024;
    if (first_in_smem) for (int i = threadIdx.x; i < n_pairs; i += blockDim.x) s_first[i] = P.first_sample[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) cmap[i] = P.class_map[i];
    for (int i = threadIdx.x; i < P.n_right + P.n_left; i += blockDim.x) {
        const int side = i >= P.n_right, d = side ? i - P.n_right : i;
        const DemuxString ds = side ? P.left[d] : P.right[d];
        s_len[side][d] = ds.len;
        uint32_t c = 0;    // class ids 1..4 were assigned in order of first appearance; re-encode to 2-bit base codes
        if (FAST) for (int k = 0; k < ds.len; ++k) c |= (uint32_t)P.fast_codes[(side ? P.n_right : 0) * 16 + d * 16 + k] << (2 * k);
        s_code[side][d] = c;
    }
    __syncthreads();
    if (FAST && threadIdx.x < 34) {
        const int side = threadIdx.x / 17, n = threadIdx.x % 17, cnt = side ? P.n_left : P.n_right;
        unsigned long long m = 0;
        for (int q = 0; q < cnt; ++q) m |= (unsigned long long)(s_len[side][q] == n) << q;
        s_lenmask[side][n] = m;
    }
    __syncthreads();
    const int thr = P.mismatch;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < P.n_reads;
         r += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t *u0 = nullptr, *u1 = nullptr, *seq = nullptr;
        int n0 = 0, n1 = 0, slen = 0;
        if (P.platform == SCAR_PLATFORM_ILLUMINA) {
            get_item(P.f0, r, u0, n0);
            get_item(P.f1, r, u1, n1);
        } else {
            get_item(P.seq, r, seq, slen);
            if (P.platform == SCAR_PLATFORM_TRUSEQ) {
                u0 = seq; n0 = slen < 6 ? slen : 6;
                n1 = slen < 6 ? slen : 6; u1 = seq + (slen - n1);
            }
        }
        uint64_t c0[2], c1[2];
        int cached0 = -1, cached1 = -1;
        uint32_t e0 = 0, x0 = 0, e1 = 0, x1 = 0;        // FAST: codes and spread invalid masks of u0 / u1
        const bool ramsden = P.platform == SCAR_PLATFORM_RAMSDEN;
        if (FAST && !ramsden) {
            uint32_t v;
            if (n0 <= 16) { encode16(u0, n0, e0, v); x0 = spread_invalid(v, n0); }
            if (n1 <= 16) { encode16(u1, n1, e1, v); x1 = spread_invalid(v, n1); }
        } else if (!ramsden) {
            class_codes(u0, n0, cmap, c0);
            class_codes(u1, n1, cmap, c1);
        }
        uint64_t ok_r = 0, ok_l = 0;   // bit d set: distinct string d accepted
        if (FAST && !ramsden) {
            // equal lengths: 2-bit Hamming decides (exact for t <= 1; for t > 1 a miss goes to the slow list);
            // other lengths go to the slow list (quick reject / bit-vector edit distance)
            // bit d of `hit`: Hamming(u, string d) <= t over the first 16 positions; strings of another length
            // are masked out with the per-length bit mask and go to the slow list
            // an all-ACGT field of the tabulated length: the accepted strings of EVERY such field were
            // worked out at context creation (same arithmetic), one load answers the side
            uint32_t hit_lo = 0, hit_hi = 0;
            uint64_t hit, eq, slow;
            if (P.lut_right && n0 == P.lut_n_right && x0 == 0u) ok_r = __ldg(P.lut_right + e0);
            else {
            const uint64_t all_r = P.n_right >= 64 ? ~0ull : ((1ull << P.n_right) - 1ull);
            for (int d = 0; d < P.n_right; ++d) {
                const uint32_t x = e0 ^ s_code[0][d];
                const uint32_t h = __popc(((x | (x >> 1)) & 0x55555555u) | x0) <= thr ? 1u : 0u;
                if (d < 32) hit_lo |= h << d; else hit_hi |= h << (d - 32);
            }
            hit = ((uint64_t)hit_hi << 32) | hit_lo;
            eq = n0 <= 16 ? s_lenmask[0][n0] : 0ull;
            ok_r = hit & eq;
            slow = all_r & (~eq | (thr > 1 ? ~hit : 0ull));
            for (uint64_t a = slow; a; a &= a - 1) {
                const int d = __ffsll((long long)a) - 1;
                if (accepts(P.right[d], P.peq, u0, n0, c0, cmap, thr, false)) ok_r |= 1ull << d;
            }
            }
            if (ok_r && P.lut_left && n1 == P.lut_n_left && x1 == 0u) ok_l = __ldg(P.lut_left + e1);
            else if (ok_r) {
                const uint64_t all_l = P.n_left >= 64 ? ~0ull : ((1ull << P.n_left) - 1ull);
                hit_lo = hit_hi = 0;
                for (int d = 0; d < P.n_left; ++d) {
                    const uint32_t x = e1 ^ s_code[1][d];
                    const uint32_t h = __popc(((x | (x >> 1)) & 0x55555555u) | x1) <= thr ? 1u : 0u;
                    if (d < 32) hit_lo |= h << d; else hit_hi |= h << (d - 32);
                }
                hit = ((uint64_t)hit_hi << 32) | hit_lo;
                eq = n1 <= 16 ? s_lenmask[1][n1] : 0ull;
                ok_l = hit & eq;
                slow = all_l & (~eq | (thr > 1 ? ~hit : 0ull));
                for (uint64_t a = slow; a; a &= a - 1) {
                    const int d = __ffsll((long long)a) - 1;
                    if (accepts(P.left[d], P.peq, u1, n1, c1, cmap, thr, false)) ok_l |= 1ull << d;
                }
            }
        } else {
            for (int d = 0; d < P.n_right; ++d) {
                const int m = s_len[0][d];
                if (ramsden) {
                    u0 = seq; n0 = slen < m ? slen : m;                       // seq[:len(right_index)]
                    if (cached0 != n0) { class_codes(u0, n0, cmap, c0); cached0 = n0; }
                }
                if (accepts(P.right[d], P.peq, u0, n0, c0, cmap, thr, true)) ok_r |= 1ull << d;
            }
            if (ok_r) {
                for (int d = 0; d < P.n_left; ++d) {
                    const int m = s_len[1][d];
                    if (ramsden) {
                        // seq[-len(left_index):] ; seq[-0:] is the whole string
                        n1 = (m == 0 || m >= slen) ? slen : m;
                        u1 = seq + (slen - n1);
                        if (cached1 != n1) { class_codes(u1, n1, cmap, c1); cached1 = n1; }
                    }
                    if (accepts(P.left[d], P.peq, u1, n1, c1, cmap, thr, true)) ok_l |= 1ull << d;
                }
            }
        }
        uint32_t best = SCAR_SAMPLE_NONE;
        if (ok_r && ok_l && !(ok_r & (ok_r - 1)) && !(ok_l & (ok_l - 1))) {
            // one acceptable string on each side (the usual case): a single matrix entry
            const int idx = (__ffsll((long long)ok_r) - 1) * P.n_left + (__ffsll((long long)ok_l) - 1);
            best = first_in_smem ? s_first[idx] : P.first_sample[idx];
        } else if (ok_r && ok_l) {
            // first manifest row whose (right, left) strings are both acceptable
            for (uint64_t a = ok_r; a; a &= a - 1) {
                const int dr = __ffsll((long long)a) - 1;
                for (uint64_t b = ok_l; b; b &= b - 1) {
                    const int dl = __ffsll((long long)b) - 1;
                    const uint32_t s = first_in_smem ? s_first[dr * P.n_left + dl] : P.first_sample[dr * P.n_left + dl];
                    best = s < best ? s : best;
                }
            }
        }
        P.samples[r] = (uint16_t)best;
    }
}

cudaError_t scar_launch_demux(const K2Params &P, int sm_count, cudaStream_t stream, int *launches)
{
    if (P.n_reads <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks = (P.n_reads + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count * 8 * 4;
    if (blocks > cap) blocks = cap;
    if (P.fast) scar_demux_kernel<true><<<(unsigned)blocks, threads, 0, stream>>>(P);
    else scar_demux_kernel<false><<<(unsigned)blocks, threads, 0, stream>>>(P);
    if (launches) ++*launches;
    return cudaGetLastError();
}

// ---- Sequence_Magic.match_maker over a batch (drop-in for the per-call helper) -------------------
// One thread per pair; classic two-row DP is replaced by the same bit-vector recurrence with the
// pattern masks built on the fly (query length <= 64), byte alphabet.
__global__ void scar_match_maker_kernel(scar_ragged q, scar_ragged u, int64_t n, int32_t *out)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint8_t *qp, *up; int m, len;
    get_item(q, r, qp, m);
    get_item(u, r, up, len);
    if (m == 0) { out[r] = len - (len - m); return; }
    uint64_t Pv = ~0ull, Mv = 0ull;
    const uint64_t top = 1ull << (m - 1);
    int score = m;
    for (int j = 0; j < len; ++j) {
        uint64_t Eq = 0ull;
        const uint8_t c = up[j];
        for (int i = 0; i < m; ++i) Eq |= (uint64_t)(qp[i] == c) << i;
        const uint64_t Xv = Eq | Mv;
        const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
        uint64_t Ph = Mv | ~(Xh | Pv);
        uint64_t Mh = Pv & Xh;
        score += (Ph & top) ? 1 : 0;
        score -= (Mh & top) ? 1 : 0;
        Ph = (Ph << 1) | 1ull;
        Mh <<= 1;
        Pv = Mh | ~(Xv | Ph);
        Mv = Ph & Xv;
    }
    out[r] = score - (len - m);
}

cudaError_t scar_launch_match_maker(const scar_ragged &q, const scar_ragged &u, int64_t n, int32_t *out,
                                    cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    const int threads = 128;
    scar_match_maker_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(q, u, n, out);
    return cudaGetLastError();
}
