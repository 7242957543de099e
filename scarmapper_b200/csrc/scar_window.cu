// scar_window.cu — K3: phasing + read filters + sliding-window junction search + junction call.
//
// Replaces, for a whole batch, the per-read body of ScarSearch.data_processing
// (reference scarmapper/INDEL_Processing.py:135-196) and SlidingWindow.sliding_window
// (scarmapper/SlidingWindow.pyx:20-171), plus the phasing scan of consensus_demultiplex
// (INDEL_Processing.py:948-975).  Not a translation: the reference compares every 10-nt read
// window with every target window; here each target side is a bucketed perfect-hash table
// "10-mer -> smallest window index i" staged in shared memory by one TMA bulk copy, a warp owns
// a read, its 32 lanes test 32 consecutive read positions per round with one shared-memory probe
// each, and the first hit in scan order is picked with a ballot.
//
//   left  search: positions p = len-20 .. 16 descending, first hit -> clj = p+10, tlj = cut - i
//   right search: positions p = 10 .. len-26 ascending,  first hit -> crj = p,    trj = cut + i
//
// Work per read: (positions scanned)/32 rounds x (1 LDS.128 + 1 LDS.64 + ~20 integer ops).
#include "scar_internal.cuh"


__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// One TMA bulk copy (cp.async.bulk, SASS UBLKCP) of the whole table blob into shared memory,
// completion signalled on an mbarrier.
__device__ __forceinline__ void stage_tables_tma(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(dst)),
            "l"(src), "r"(bytes), "r"(smem_u32(bar))
            : "memory");
    }
    // every thread waits for phase 0
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(0u)
            : "memory");
    }
}

// probe one side's table: returns the stored entry ((key|TAG) << 11 | i) or 0
template <bool SMEM>
__device__ __forceinline__ uint32_t probe(const uint2 *tab, uint32_t key, uint32_t mult, uint32_t shift)
{
    const uint32_t b = (key * mult) >> shift;
    uint2 e;
    if (SMEM) e = tab[b]; else e = __ldg(tab + b);
    const uint32_t want = key | SCAR_KEY_TAG;
    uint32_t hit = 0;
    if ((e.x >> 11) == want) hit = e.x;
    if ((e.y >> 11) == want) hit = e.y;
    return hit;
}

template <bool SMEM, bool HR>
__global__ void __launch_bounds__(256) scar_window_kernel(const K3Params P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [mbarrier 16 B][tables (SMEM only)][per-warp staging rows of uint4]
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    const uint2 *tab = P.tables;
    uint32_t off = 16;
    if (SMEM) {
        uint2 *stab = reinterpret_cast<uint2 *>(smem_raw + off);
        stage_tables_tma(stab, P.tables, P.table_bytes, bar);
        tab = stab;
        off += P.table_bytes;
    }
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    uint4 *row = reinterpret_cast<uint4 *>(smem_raw + off) + (size_t)warp * P.row_entries;
    const uint32_t FULL = 0xFFFFFFFFu;
    const int W = P.W;

    for (int64_t r = (int64_t)blockIdx.x * warps_per_cta + warp; r < P.n_reads;
         r += (int64_t)gridDim.x * warps_per_cta) {
        const uint32_t sample = P.samples[r];
        __align__(16) scar_record rec;
        rec.status = SCAR_ST_UNASSIGNED; rec.flags = 0; rec.sample = (uint16_t)sample;
        rec.clj = rec.crj = rec.tlj = rec.trj = 0; rec.hr_n = 0;
        rec.phase_r1 = SCAR_PHASE_NA; rec.phase_r2 = SCAR_PHASE_NA;
        if (sample >= (uint32_t)P.n_samples) {          // unidentified (INDEL_Processing.py:1194-1198)
            rec.sample = SCAR_SAMPLE_NONE;
            if (lane == 0) *reinterpret_cast<uint4 *>(P.records + r) = *reinterpret_cast<uint4 *>(&rec);
            continue;
        }
        const int len = P.lens[r];
        const ScarTargetMeta T = P.targets[P.sample_target[sample]];
        const uint64_t *crow = P.cells + r * W;

        // ---- stage the row: entry j = {codes[j], codes[j+1], valid[j]|valid[j+1]<<16, N[j]|N[j+1]<<16}
        __syncwarp();
        int n_count = 0;
        for (int j = lane; j < P.row_entries; j += 32) {
            const uint64_t c0 = j < W ? crow[j] : 0ull;
            const uint64_t c1 = j + 1 < W ? crow[j + 1] : 0ull;
            uint4 e;
            e.x = (uint32_t)c0; e.y = (uint32_t)c1;
            e.z = ((uint32_t)(c0 >> 32) & 0xFFFFu) | (((uint32_t)(c1 >> 32) & 0xFFFFu) << 16);
            e.w = (uint32_t)(c0 >> 48) | ((uint32_t)(c1 >> 48) << 16);
            row[j] = e;
            n_count += __popc((uint32_t)(c0 >> 48));
        }
        n_count = __reduce_add_sync(FULL, n_count);
        __syncwarp();

        // ---- phasing (INDEL_Processing.py:948-975): lane k owns phase k of the read's locus
        if (P.do_phasing && T.n_phase > 0) {
            bool m1 = false, m2 = false;
            if (lane < T.n_phase) {
                const ScarPhase ph = P.phases[T.phase_off + lane];
                if (ph.r1_len != 0) {                       // empty R1 string: phase skipped (:954-957)
                    if (ph.r1_len <= 16 && len >= ph.r1_len) {
                        const uint4 e = row[0];
                        const uint32_t vm = (1u << ph.r1_len) - 1u;
                        const uint32_t cm = ph.r1_len == 16 ? 0xFFFFFFFFu : ((1u << (2 * ph.r1_len)) - 1u);
                        m1 = ((e.z & vm) == vm) && ((e.x & cm) == ph.r1_code);
                    }
                    if (ph.r2_len != 0 && ph.r2_len <= 16 && len >= ph.r2_len) {
                        const int p = len - ph.r2_len;
                        const uint4 e = row[p >> 4];
                        const int o = p & 15;
                        const uint32_t vm = (1u << ph.r2_len) - 1u;
                        const uint32_t cm = ph.r2_len == 16 ? 0xFFFFFFFFu : ((1u << (2 * ph.r2_len)) - 1u);
                        m2 = (((e.z >> o) & vm) == vm) && ((__funnelshift_r(e.x, e.y, 2 * o) & cm) == ph.r2_code);
                    }
                }
            }
            const uint32_t b1 = __ballot_sync(FULL, m1), b2 = __ballot_sync(FULL, m2);
            rec.phase_r1 = b1 ? (int8_t)(__ffs(b1) - 1) : (int8_t)SCAR_PHASE_NONE;
            rec.phase_r2 = b2 ? (int8_t)(__ffs(b2) - 1) : (int8_t)SCAR_PHASE_NONE;
        }

        // ---- read filters (INDEL_Processing.py:147-154); Python int/int is an IEEE double divide
        if (len == 0 || (double)n_count / (double)len > P.n_limit || len <= P.min_length) {
            rec.status = SCAR_ST_FILTERED;
            if (lane == 0) *reinterpret_cast<uint4 *>(P.records + r) = *reinterpret_cast<uint4 *>(&rec);
            continue;
        }

        int clj = 0, crj = 0, tlj = T.cutsite, trj = T.cutsite;
        bool cutwindow_exit = false;

        // ---- left junction (SlidingWindow.pyx:49-69): p = len-20 .. 16, descending
        {
            const uint2 *ltab = tab + T.left_off;
            for (int pstart = len - 20; pstart >= 16; pstart -= 32) {
                const int p = pstart - lane;
                uint32_t hit = 0;
                if (p >= 16) {
                    const uint4 e = row[p >> 4];
                    const int o = p & 15;
                    if (((e.z >> o) & 0x3FFu) == 0x3FFu) {
                        const uint32_t key = __funnelshift_r(e.x, e.y, 2 * o) & SCAR_KEY_MASK;
                        hit = probe<SMEM>(ltab, key, T.left_mult, T.left_shift);
                    }
                }
                const uint32_t bal = __ballot_sync(FULL, hit != 0);
                if (bal) {
                    const int src = __ffs(bal) - 1;            // lowest lane = highest position
                    const uint32_t h = __shfl_sync(FULL, hit, src);
                    if (((h >> 11) & SCAR_KEY_MASK) == T.cutwindow_key) cutwindow_exit = true;  // :56-60
                    clj = pstart - src + 10;
                    tlj = T.cutsite - (int)(h & 0x7FFu);
                    break;
                }
            }
        }
        if (cutwindow_exit) {
            rec.status = SCAR_ST_NO_CUT; rec.flags = SCAR_FL_CUTWINDOW;
            rec.tlj = rec.trj = (uint16_t)T.cutsite;
            if (lane == 0) *reinterpret_cast<uint4 *>(P.records + r) = *reinterpret_cast<uint4 *>(&rec);
            continue;
        }

        // ---- right junction (SlidingWindow.pyx:79-93): p = 10 .. len-26, ascending
        {
            const uint2 *rtab = tab + T.right_off;
            const int plast = len - 26;
            for (int pstart = 10; pstart <= plast; pstart += 32) {
                const int p = pstart + lane;
                uint32_t hit = 0;
                if (p <= plast) {
                    const uint4 e = row[p >> 4];
                    const int o = p & 15;
                    if (((e.z >> o) & 0x3FFu) == 0x3FFu) {
                        const uint32_t key = __funnelshift_r(e.x, e.y, 2 * o) & SCAR_KEY_MASK;
                        hit = probe<SMEM>(rtab, key, T.right_mult, T.right_shift);
                    }
                }
                const uint32_t bal = __ballot_sync(FULL, hit != 0);
                if (bal) {
                    const int src = __ffs(bal) - 1;
                    const uint32_t h = __shfl_sync(FULL, hit, src);
                    crj = pstart + src;
                    trj = T.cutsite + (int)(h & 0x7FFu);
                    break;
                }
            }
        }

        rec.clj = (uint16_t)clj; rec.crj = (uint16_t)crj; rec.tlj = (uint16_t)tlj; rec.trj = (uint16_t)trj;

        if (clj < 1 && crj < 1) {                            // :96-98
            rec.status = SCAR_ST_NO_JUNCTION;
            if (lane == 0) *reinterpret_cast<uint4 *>(P.records + r) = *reinterpret_cast<uint4 *>(&rec);
            continue;
        }

        // ---- HR donor scan (:102-117): occurrences at p = 25 .. len-26-h
        if (HR) {
            const int h = P.hr_len;
            const int plast = len - 26 - h;
            const uint64_t cm = h == 32 ? ~0ull : ((1ull << (2 * h)) - 1ull);
            const uint64_t vm = h == 32 ? 0xFFFFFFFFull : ((1ull << h) - 1ull);
            int n = 0;
            for (int pstart = 25; pstart <= plast; pstart += 32) {
                const int p = pstart + lane;
                bool m = false;
                if (p <= plast) {
                    const int j = p >> 4, o = p & 15;
                    const uint4 e0 = row[j], e1 = row[j + 1];      // cells j, j+1 | j+1, j+2
                    const uint32_t lo = __funnelshift_r(e0.x, e0.y, 2 * o);
                    const uint32_t hi = __funnelshift_r(e1.x, e1.y, 2 * o);
                    const uint64_t code = ((uint64_t)hi << 32 | lo) & cm;
                    const uint64_t v48 = (uint64_t)e0.z | ((uint64_t)(e1.z >> 16) << 32);
                    m = (((v48 >> o) & vm) == vm) && code == P.hr_code;
                }
                n += __popc(__ballot_sync(FULL, m));
            }
            rec.hr_n = (uint16_t)n;
            if (n) rec.flags |= SCAR_FL_HR;
        }

        // ---- junction call (:134-171)
        bool cut_found = false;
        if (0 < clj && clj < crj) {
            // literal 'N' anywhere in read[clj:crj] drops the read with no counter (:140-141)
            bool has_n = false;
            for (int j = lane; j < W; j += 32) {
                const int lo = max(clj, 16 * j) - 16 * j, hi = min(crj, 16 * j + 16) - 16 * j;
                if (hi > lo) {
                    const uint32_t m = ((1u << (hi - lo)) - 1u) << lo;
                    has_n |= ((row[j].w & 0xFFFFu) & m) != 0;
                }
            }
            if (__any_sync(FULL, has_n)) {
                rec.status = SCAR_ST_N_INSERTION;
                if (lane == 0) *reinterpret_cast<uint4 *>(P.records + r) = *reinterpret_cast<uint4 *>(&rec);
                continue;
            }
            cut_found = true; rec.flags |= SCAR_FL_INSERTION;
        }
        if (tlj < T.cutsite) { cut_found = true; rec.flags |= SCAR_FL_LEFT_DEL; }
        if (trj > T.cutsite) { cut_found = true; rec.flags |= SCAR_FL_RIGHT_DEL; }
        if (clj > crj && crj > 0) { cut_found = true; rec.flags |= SCAR_FL_MICROHOM; }
        rec.status = cut_found ? SCAR_ST_SCAR : SCAR_ST_NO_CUT;
        if (lane == 0) *reinterpret_cast<uint4 *>(P.records + r) = *reinterpret_cast<uint4 *>(&rec);
    }
}

// host-side launcher (called from scar_api.cu)
cudaError_t scar_launch_window(const K3Params &P, bool smem_tables, int sm_count, cudaStream_t stream,
                               int *launches)
{
    const int threads = 256;
    const size_t row_bytes = (size_t)(threads / 32) * P.row_entries * sizeof(uint4);
    const size_t smem = 16 + (smem_tables ? P.table_bytes : 0) + row_bytes;
    const bool hr = P.hr_len > 0;
    void (*kern)(const K3Params) =
        smem_tables ? (hr ? scar_window_kernel<true, true> : scar_window_kernel<true, false>)
                    : (hr ? scar_window_kernel<false, true> : scar_window_kernel<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    // persistent grid: a whole number of resident CTAs per SM, never more than the work needs
    int64_t want = (P.n_reads + (threads / 32) - 1) / (threads / 32);
    int64_t grid = (int64_t)sm_count * per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, threads, smem, stream>>>(P);
    if (launches) ++*launches;
    return cudaGetLastError();
}
