// scar_table.cu — K4: per-sample counters, phase histograms and the scar frequency table.
//
// Replaces stage 4 of ScarSearch.data_processing (reference scarmapper/INDEL_Processing.py:186-196:
// results_freq_dict[key] = [count, first sub_list]) and the summary_data / read_count_dict /
// phase_count bookkeeping (:94,147-169,948-975,1173-1198) for a whole batch.
//
// The reference key is the text "ldel|rdel|ins|mh|hr".  For one sample (= one target) ldel and rdel
// are determined by (tlj, trj), and ins / mh are read[clj:crj] / read[crj:clj], so the key is
// (sample, tlj, trj, kind, segment bytes, hr).  It is identified by two independent 64-bit hashes
// (128 bits; both must agree).  The table is open-addressed in HBM; a slot keeps the count and the
// smallest global read index (the reference stores the FIRST read's record and orders ties by
// first-seen rank).
#include "scar_table_core.cuh"

static __device__ __forceinline__ ScarTableRef table_ref(const K4Params &P)
{
    ScarTableRef T; T.slots = P.slots; T.slot_mask = P.slot_mask; T.max_probe = P.max_probe; T.status = P.status;
    T.claimed = P.claimed; T.n_entries = P.n_entries;
    return T;
}
static __device__ __forceinline__ ScarAccRef acc_ref(const K4Params &P)
{
    ScarAccRef A; A.counters = P.counters; A.phase_hist = P.phase_hist; A.unassigned = P.unassigned;
    A.n_samples = P.n_samples; A.pb = P.phase_bins; A.joint = P.phase_joint;
    return A;
}

#ifndef SCAR_K4_CTAS
#define SCAR_K4_CTAS 6
#endif
__global__ void __launch_bounds__(256, SCAR_K4_CTAS) scar_aggregate_kernel(const K4Params P)
{
    extern __shared__ __align__(8) unsigned int s_hist[];
    const ScarTableRef T = table_ref(P);
    const ScarAccRef A = acc_ref(P);
    const int per_sample = acc_words_per_sample(A.pb, A.joint);
    __shared__ unsigned int s_unassigned;
    if (P.use_smem) {
        for (int i = threadIdx.x; i < P.n_samples * per_sample; i += blockDim.x) s_hist[i] = 0;
    }
    if (threadIdx.x == 0) s_unassigned = 0;
    __syncthreads();
    int iter = 0;
    // the trip count is the same for every thread of a CTA (the epoch flush below is a CTA barrier)
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < P.n_reads; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = base + threadIdx.x;
        if (P.use_smem && ++iter > K4_EPOCH_ITERS) {
            __syncthreads();
            acc_flush(A, s_hist, per_sample);
            __syncthreads();
            iter = 1;
        }
        const bool in_range = r < P.n_reads;
        __align__(16) scar_record rec;
        if (in_range) *reinterpret_cast<uint4 *>(&rec) = __ldg(reinterpret_cast<const uint4 *>(P.records + r));
        else { rec.status = 0xFF; rec.flags = 0; rec.sample = 0; rec.hr_n = 0; rec.phase_r1 = rec.phase_r2 = SCAR_PHASE_NA; }

        // ---- counters (summary_data, read_count_dict, phase_count)
        if (in_range) {
            if (rec.status == SCAR_ST_UNASSIGNED) atomicAdd(&s_unassigned, 1u);
            else if (P.use_smem) acc_add_smem(s_hist, per_sample, A.pb, A.joint, rec);
            else acc_add_global(A, rec);
        }

        // ---- scar table: every SCAR read inserts its own key.  (Combining equal keys inside a warp with
        // __match_any_sync / per-group shuffles was measured slower: most keys of a warp are distinct and
        // every distinct group is a separate divergent pass; the L2 atomics of equal keys coalesce anyway.)
        bool claimed = false;
        uint32_t slot = 0;
        if (in_range && rec.status == SCAR_ST_SCAR) {
            const uint8_t *raw; int rawlen;
            get_item(P.seq, r, raw, rawlen);
            uint64_t h1, h2;
            key_hashes(rec, raw, rawlen, P.platform, h1, h2, P.weak_keys != 0);
            claimed = table_insert(T, h1, h2, 1u, P.first_index + (uint64_t)r, rec.sample, slot);
        }
        table_note_claims(T, claimed, slot);          // (the loop's trip count is uniform: the warp is converged)
    }
    __syncthreads();
    if (P.use_smem) acc_flush(A, s_hist, per_sample);
    if (threadIdx.x == 0 && s_unassigned) atomicAdd(P.unassigned, (unsigned long long)s_unassigned);
}

// ---- verification: every SCAR read's full key equals the key of its slot's first read ------------

__global__ void __launch_bounds__(256) scar_verify_kernel(const K4VerifyParams P)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < P.n_reads;
         r += (int64_t)gridDim.x * blockDim.x) {
        __align__(16) scar_record rec;
        *reinterpret_cast<uint4 *>(&rec) = __ldg(reinterpret_cast<const uint4 *>(P.records + r));
        if (rec.status != SCAR_ST_SCAR) continue;
        const uint8_t *raw; int rawlen;
        get_item(P.seq, r, raw, rawlen);
        uint64_t h1, h2;
        key_hashes(rec, raw, rawlen, P.platform, h1, h2, P.weak_keys != 0);
        uint64_t slot = h1 & P.slot_mask;
        bool found = false;
        for (uint32_t probe = 0; probe <= P.max_probe; ++probe) {
            const ScarTableSlot s = P.slots[slot];
            if (s.tag == 0) break;
            if (s.tag == make_tag(h1, h2)) {
                found = true;
                const uint64_t first = ~s.inv_first;
                // the representative is compared only when it lies in this resident batch
                if (first >= P.first_index && first < P.first_index + (uint64_t)P.n_reads) {
                    const int64_t q = (int64_t)(first - P.first_index);
                    __align__(16) scar_record rr;
                    *reinterpret_cast<uint4 *>(&rr) = __ldg(reinterpret_cast<const uint4 *>(P.records + q));
                    const uint8_t *raw2; int rawlen2;
                    get_item(P.seq, q, raw2, rawlen2);
                    bool eq = rr.status == SCAR_ST_SCAR && rr.sample == rec.sample && rr.tlj == rec.tlj && rr.trj == rec.trj &&
                              ((rr.flags ^ rec.flags) & (SCAR_FL_INSERTION | SCAR_FL_MICROHOM | SCAR_FL_HR)) == 0;
                    int lo = 0, hi = 0, lo2 = 0, hi2 = 0;
                    if (rec.flags & SCAR_FL_INSERTION) { lo = rec.clj; hi = rec.crj; } else if (rec.flags & SCAR_FL_MICROHOM) { lo = rec.crj; hi = rec.clj; }
                    if (rr.flags & SCAR_FL_INSERTION) { lo2 = rr.clj; hi2 = rr.crj; } else if (rr.flags & SCAR_FL_MICROHOM) { lo2 = rr.crj; hi2 = rr.clj; }
                    eq = eq && (hi - lo) == (hi2 - lo2);
                    for (int k = 0; eq && k < hi - lo; ++k)
                        eq = stored_byte(raw, rawlen, P.platform, lo + k) == stored_byte(raw2, rawlen2, P.platform, lo2 + k);
                    if (!eq) atomicAdd(P.n_collisions, 1ull);
                }
                break;
            }
            slot = (slot + 1) & P.slot_mask;
        }
        if (!found) atomicAdd(P.n_collisions, 1ull);
    }
}

// ---- export and merge: both walk the claimed-slot list, not the table ---------------------------------
__device__ __forceinline__ scar_entry entry_of(const ScarTableSlot &s)
{
    scar_entry e; e.h1 = (uint64_t)s.tag; e.h2 = (uint64_t)(s.tag >> 64); e.first_idx = ~s.inv_first;
    e.count = s.count; e.sample = (uint16_t)s.sample; e.pad = 0;
    return e;
}

__global__ void scar_export_kernel(const ScarTableSlot *slots, const uint32_t *claimed, const unsigned long long *n_entries,
                                   scar_entry *out, uint64_t out_capacity)
{
    const uint64_t n = *n_entries;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n && i < out_capacity;
         i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = entry_of(slots[claimed[i]]);
}

// empty the claimed slots (the count is reset by the caller, stream-ordered after this kernel)
__global__ void scar_clear_kernel(ScarTableSlot *slots, const uint32_t *claimed, const unsigned long long *n_entries)
{
    const uint64_t n = *n_entries;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 *p = reinterpret_cast<uint4 *>(slots + claimed[i]);
        p[0] = make_uint4(0u, 0u, 0u, 0u); p[1] = make_uint4(0u, 0u, 0u, 0u);
    }
}

__global__ void scar_merge_kernel(K4Params P, const scar_entry *in, uint64_t n_in)
{
    const ScarTableRef T = table_ref(P);
    const uint64_t n_round = (n_in + 31) & ~(uint64_t)31;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += (uint64_t)gridDim.x * blockDim.x) {
        bool claimed = false;
        uint32_t slot = 0;
        if (i < n_in) {
            const scar_entry e = in[i];
            claimed = table_insert(T, e.h1, e.h2, e.count, e.first_idx, e.sample, slot);
        }
        table_note_claims(T, claimed, slot);
    }
}

// ---- owner-partitioned export / fold (multi-GPU merge whose cost does not grow with the GPU count) -----
// The table is cut by key ownership: owner = (h2 >> 40) % n_parts.  Region o of the output (capacity
// `region` entries, entry 0 is a header whose h1 holds the count) receives this rank's entries owned by
// rank o; the regions are exchanged with one all-to-all and every rank folds what it receives.  The
// counts travel in band, so the whole merge is stream-ordered: no host synchronisation.
__global__ void scar_export_parts_kernel(const ScarTableSlot *slots, const uint32_t *claimed, const unsigned long long *n_entries,
                                         scar_entry *out, uint32_t n_parts, uint64_t region, int32_t *status)
{
    const uint64_t n = *n_entries, n_round = (n + 31) & ~(uint64_t)31;
    const int lane = threadIdx.x & 31;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const bool used = i < n;
        ScarTableSlot s;
        s.tag = 0;
        if (used) s = slots[claimed[i]];
        const uint64_t h2 = (uint64_t)(s.tag >> 64);
        const uint32_t owner = used ? (uint32_t)(h2 >> 40) % n_parts : 0xFFFFFFFFu;
        // one atomic per (warp, owner): the region cursors are only n_parts addresses
        unsigned long long pos = 0;
        for (uint32_t o = 0; o < n_parts; ++o) {
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, owner == o);
            if (!bal) continue;
            unsigned long long base = 0;
            if (lane == __ffs(bal) - 1)
                base = atomicAdd(reinterpret_cast<unsigned long long *>(&out[(uint64_t)o * region].h1),
                                 (unsigned long long)__popc(bal));
            base = __shfl_sync(0xFFFFFFFFu, base, __ffs(bal) - 1);
            if (owner == o) pos = 1ull + base + __popc(bal & ((1u << lane) - 1u));
        }
        if (used) {
            if (pos < region) out[(uint64_t)owner * region + pos] = entry_of(s);
            else atomicOr(status, 4);
        }
    }
}

// The same export fused with the exchange: every entry is stored straight into the receive buffer of its
// OWNER GPU through NVLink peer memory (peers.base[o] is rank o's receive buffer mapped into this process;
// this rank writes region `my_rank` of it), so no separate all-to-all pass runs and the transfer overlaps
// the table walk.  The region cursors stay local; a second tiny kernel publishes the counts as headers.
struct ScarPeers { scar_entry *base[SCAR_MAX_PEERS]; };

__global__ void scar_export_p2p_kernel(const ScarTableSlot *slots, const uint32_t *claimed, const unsigned long long *n_entries,
                                       ScarPeers peers, uint32_t n_parts, uint32_t my_rank, uint64_t region,
                                       unsigned long long *cursors)
{
    const uint64_t n = *n_entries, n_round = (n + 31) & ~(uint64_t)31;
    const int lane = threadIdx.x & 31;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const bool used = i < n;
        ScarTableSlot s;
        s.tag = 0;
        if (used) s = slots[claimed[i]];
        const uint64_t h2 = (uint64_t)(s.tag >> 64);
        const uint32_t owner = used ? (uint32_t)(h2 >> 40) % n_parts : 0xFFFFFFFFu;
        unsigned long long pos = 0;
        for (uint32_t o = 0; o < n_parts; ++o) {       // one atomic per (warp, owner)
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, owner == o);
            if (!bal) continue;
            unsigned long long base = 0;
            if (lane == __ffs(bal) - 1) base = atomicAdd(cursors + o, (unsigned long long)__popc(bal));
            base = __shfl_sync(0xFFFFFFFFu, base, __ffs(bal) - 1);
            if (owner == o) pos = 1ull + base + __popc(bal & ((1u << lane) - 1u));
        }
        if (used && pos < region) {                    // (an overflow is reported by the header kernel)
            const scar_entry e = entry_of(s);
            uint4 *dst = reinterpret_cast<uint4 *>(peers.base[owner] + (uint64_t)my_rank * region + pos);
            const uint4 *src = reinterpret_cast<const uint4 *>(&e);
            dst[0] = src[0]; dst[1] = src[1];          // two 16-byte stores over NVLink
        }
    }
}

__global__ void scar_export_p2p_headers_kernel(ScarPeers peers, uint32_t n_parts, uint32_t my_rank, uint64_t region,
                                               unsigned long long *cursors, int32_t *status)
{
    const uint32_t o = threadIdx.x;
    if (o >= n_parts) return;
    unsigned long long n = cursors[o];
    cursors[o] = 0;                                    // ready for the next export
    if (n > region - 1) { atomicOr(status, 4); n = region - 1; }
    scar_entry h; h.h1 = n; h.h2 = 0; h.first_idx = 0; h.count = 0; h.sample = 0; h.pad = 0;
    uint4 *dst = reinterpret_cast<uint4 *>(peers.base[o] + (uint64_t)my_rank * region);
    const uint4 *src = reinterpret_cast<const uint4 *>(&h);
    dst[0] = src[0]; dst[1] = src[1];
}

cudaError_t scar_launch_export_p2p(const ScarTableSlot *slots, const uint32_t *claimed, const unsigned long long *n_entries,
                                   scar_entry *const *peer_base, uint32_t n_parts, uint32_t my_rank, uint64_t region,
                                   unsigned long long *cursors, int32_t *status, int sm_count, cudaStream_t stream)
{
    ScarPeers peers;
    for (uint32_t o = 0; o < SCAR_MAX_PEERS; ++o) peers.base[o] = o < n_parts ? peer_base[o] : nullptr;
    scar_export_p2p_kernel<<<(unsigned)(sm_count * 4), 256, 0, stream>>>(slots, claimed, n_entries, peers, n_parts, my_rank, region, cursors);
    scar_export_p2p_headers_kernel<<<1, 32, 0, stream>>>(peers, n_parts, my_rank, region, cursors, status);
    return cudaGetLastError();
}

__global__ void scar_merge_parts_kernel(K4Params P, const scar_entry *in, uint32_t n_parts, uint64_t region)
{
    const ScarTableRef T = table_ref(P);
    for (uint32_t g = 0; g < n_parts; ++g) {
        const scar_entry *reg = in + (uint64_t)g * region;
        uint64_t cnt = reg->h1;
        if (cnt > region - 1) cnt = region - 1;
        const uint64_t n_round = (cnt + 31) & ~(uint64_t)31;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round;
             i += (uint64_t)gridDim.x * blockDim.x) {
            bool claimed = false;
            uint32_t slot = 0;
            if (i < cnt) {
                const scar_entry e = reg[1 + i];
                claimed = table_insert(T, e.h1, e.h2, e.count, e.first_idx, e.sample, slot);
            }
            table_note_claims(T, claimed, slot);
        }
    }
}

cudaError_t scar_launch_export_parts(const ScarTableSlot *slots, const uint32_t *claimed, const unsigned long long *n_entries,
                                     scar_entry *out, uint32_t n_parts, uint64_t region, int32_t *status, int sm_count,
                                     cudaStream_t stream)
{
    scar_export_parts_kernel<<<(unsigned)(sm_count * 4), 256, 0, stream>>>(slots, claimed, n_entries, out, n_parts, region, status);
    return cudaGetLastError();
}

cudaError_t scar_launch_merge_parts(const K4Params &P, const scar_entry *in, uint32_t n_parts, uint64_t region,
                                    int sm_count, cudaStream_t stream)
{
    scar_merge_parts_kernel<<<(unsigned)(sm_count * 4), 256, 0, stream>>>(P, in, n_parts, region);
    return cudaGetLastError();
}

cudaError_t scar_launch_aggregate(K4Params P, int sm_count, cudaStream_t stream, int *launches)
{
    if (P.n_reads <= 0) return cudaSuccess;
    const int threads = 256;
    const size_t per_sample = (size_t)acc_words_per_sample(P.phase_bins, P.phase_joint) * sizeof(unsigned int);
    size_t smem = (size_t)P.n_samples * per_sample;
    P.use_smem = smem <= 96 * 1024;
    if (!P.use_smem) smem = 0;
    cudaError_t e = cudaFuncSetAttribute(scar_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(96 * 1024));
    if (e != cudaSuccess) return e;
    int64_t blocks = (P.n_reads + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count * SCAR_K4_CTAS;
    if (blocks > cap) blocks = cap;
    scar_aggregate_kernel<<<(unsigned)blocks, threads, smem, stream>>>(P);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t scar_launch_verify(const K4VerifyParams &P, int sm_count, cudaStream_t stream)
{
    if (P.n_reads <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks = (P.n_reads + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    scar_verify_kernel<<<(unsigned)blocks, threads, 0, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t scar_launch_export(const ScarTableSlot *slots, const uint32_t *claimed, const unsigned long long *n_entries,
                               scar_entry *out, uint64_t out_capacity, int sm_count, cudaStream_t stream)
{
    scar_export_kernel<<<(unsigned)(sm_count * 4), 256, 0, stream>>>(slots, claimed, n_entries, out, out_capacity);
    return cudaGetLastError();
}

cudaError_t scar_launch_clear(ScarTableSlot *slots, const uint32_t *claimed, unsigned long long *n_entries, int sm_count,
                              cudaStream_t stream)
{
    scar_clear_kernel<<<(unsigned)(sm_count * 4), 256, 0, stream>>>(slots, claimed, n_entries);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaMemsetAsync(n_entries, 0, sizeof(unsigned long long), stream);
}

cudaError_t scar_launch_merge(const K4Params &P, const scar_entry *in, uint64_t n_in, int sm_count, cudaStream_t stream)
{
    if (n_in == 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks = (int64_t)((n_in + threads - 1) / threads);
    const int64_t cap = (int64_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    scar_merge_kernel<<<(unsigned)blocks, threads, 0, stream>>>(P, in, n_in);
    return cudaGetLastError();
}
