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


__global__ void __launch_bounds__(256) scar_demux_kernel(const K2Params P)
{
    __shared__ uint8_t cmap[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) cmap[i] = P.class_map[i];
    __syncthreads();
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
        if (P.platform != SCAR_PLATFORM_RAMSDEN) {
            class_codes(u0, n0, cmap, c0);
            class_codes(u1, n1, cmap, c1);
        }
        uint64_t ok_r = 0, ok_l = 0;   // bit d set: distinct string d accepted
        for (int d = 0; d < P.n_right; ++d) {
            const DemuxString ds = P.right[d];
            if (P.platform == SCAR_PLATFORM_RAMSDEN) {
                u0 = seq; n0 = slen < ds.len ? slen : ds.len;             // seq[:len(right_index)]
                if (cached0 != n0) { class_codes(u0, n0, cmap, c0); cached0 = n0; }
            }
            if (accepts(ds, P.peq, u0, n0, c0, cmap, P.mismatch)) ok_r |= 1ull << d;
        }
        uint32_t best = SCAR_SAMPLE_NONE;
        if (ok_r) {
            for (int d = 0; d < P.n_left; ++d) {
                const DemuxString ds = P.left[d];
                if (P.platform == SCAR_PLATFORM_RAMSDEN) {
                    // seq[-len(left_index):] ; seq[-0:] is the whole string
                    n1 = (ds.len == 0 || ds.len >= slen) ? slen : ds.len;
                    u1 = seq + (slen - n1);
                    if (cached1 != n1) { class_codes(u1, n1, cmap, c1); cached1 = n1; }
                }
                if (accepts(ds, P.peq, u1, n1, c1, cmap, P.mismatch)) ok_l |= 1ull << d;
            }
            // first manifest row whose (right, left) strings are both acceptable
            for (uint64_t a = ok_r; a; a &= a - 1) {
                const int dr = __ffsll((long long)a) - 1;
                for (uint64_t b = ok_l; b; b &= b - 1) {
                    const int dl = __ffsll((long long)b) - 1;
                    const uint32_t s = P.first_sample[dr * P.n_left + dl];
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
    scar_demux_kernel<<<(unsigned)blocks, threads, 0, stream>>>(P);
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
