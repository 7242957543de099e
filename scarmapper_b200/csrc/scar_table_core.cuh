// scar_table_core.cuh — device code shared by K4 (scar_table.cu) and the fused classify+aggregate kernel
// (scar_window.cu): the identity of a scar key, the open-addressed table insert and the per-sample accumulators.
//
// The reference key is the text "ldel|rdel|ins|mh|hr" (scarmapper/INDEL_Processing.py:186-196).  For one sample
// (= one target) ldel and rdel are determined by (tlj, trj), and ins / mh are read[clj:crj] / read[crj:clj], so the
// key is (sample, tlj, trj, kind, segment, hr).  It is identified by two independent 64-bit hashes (128 bits; both
// must agree).  The segment is hashed in one of two ways, chosen by its CONTENT (so equal keys always choose alike):
//   all bases A/C/G/T : over its 2-bit codes, 32 bases per step - computable from the packed cells in shared memory
//                       (fused kernel) or from the raw bytes (K4, verification), with identical results;
//   otherwise         : over its raw bytes, 8 per step (rare: the reference drops reads with 'N' in the insertion).
#pragma once
#include "scar_internal.cuh"

__device__ __forceinline__ uint8_t comp_byte(uint8_t c)
{
    // Sequence_Magic.rcomp's translate table (Valkyries/Sequence_Magic.py:97-118); unmapped bytes pass
    switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'M': return 'K'; case 'R': return 'Y'; case 'Y': return 'R'; case 'K': return 'M';
    case 'V': return 'B'; case 'H': return 'D'; case 'D': return 'H'; case 'B': return 'V';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    case 'm': return 'k'; case 'r': return 'y'; case 'y': return 'r'; case 'k': return 'm';
    case 'v': return 'b'; case 'h': return 'd'; case 'd': return 'h'; case 'b': return 'v';
    default: return c;
    }
}

// byte q of the STORED sequence (after the platform transform) of a raw read
__device__ __forceinline__ uint8_t stored_byte(const uint8_t *raw, int rawlen, int platform, int q)
{
    if (platform == SCAR_PLATFORM_TRUSEQ) return raw[q + 6];
    if (platform == SCAR_PLATFORM_RAMSDEN) return comp_byte(raw[rawlen - 1 - q]);
    return raw[q];
}

// ---- key identity -----------------------------------------------------------------------------------
struct KeySegment { int lo, hi; uint64_t hdr; };

__device__ __forceinline__ KeySegment key_segment(const scar_record &rec)
{
    KeySegment k; k.lo = 0; k.hi = 0;
    uint64_t kind = 0;
    if (rec.flags & SCAR_FL_INSERTION) { k.lo = rec.clj; k.hi = rec.crj; kind = 1; }
    else if (rec.flags & SCAR_FL_MICROHOM) { k.lo = rec.crj; k.hi = rec.clj; kind = 2; }
    k.hdr = (uint64_t)rec.sample | ((uint64_t)rec.tlj << 16) | ((uint64_t)rec.trj << 32) | (kind << 48) |
            ((uint64_t)((rec.flags & SCAR_FL_HR) ? 1 : 0) << 50) | ((uint64_t)(k.hi - k.lo) << 51);   // bit 62: raw-byte scheme
    return k;
}

__device__ __forceinline__ void key_init(uint64_t hdr, uint64_t &h1, uint64_t &h2)
{
    h1 = scar_mix1(hdr ^ 0x9E3779B97F4A7C15ull);
    h2 = scar_mix2(hdr + 0xD6E8FEB86659FD93ull);
}
__device__ __forceinline__ void key_step(uint64_t w, uint64_t &h1, uint64_t &h2)
{
    h1 = scar_mix1(h1 ^ w);
    h2 = scar_mix2(h2 + w * 0xA24BAED4963EE407ull + 1ull);
}
__device__ __forceinline__ void key_finish(uint64_t &h1, uint64_t &h2)
{
    if (h1 == 0) h1 = 1;
    if (h2 == 0) h2 = 1;
}

// raw-byte scheme: eight stored bytes per step
__device__ __forceinline__ void key_hashes_bytes(const KeySegment &k, const uint8_t *raw, int rawlen, int platform,
                                                 uint64_t &h1, uint64_t &h2)
{
    key_init(k.hdr | (1ull << 62), h1, h2);
    for (int q = k.lo; q < k.hi; q += 8) {
        uint64_t w = 0;
        const int n = min(8, k.hi - q);
        for (int i = 0; i < n; ++i) w |= (uint64_t)stored_byte(raw, rawlen, platform, q + i) << (8 * i);
        key_step(w, h1, h2);
    }
    key_finish(h1, h2);
}

// the same for platforms whose stored sequence is a plain slice of the raw read (`stored` = first stored byte): the
// fused kernel's rare path, kept out of line so that the hot path stays small
static __device__ __noinline__ void key_hashes_bytes_fwd(uint64_t hdr, int lo, int hi, const uint8_t *stored, uint64_t *h)
{
    uint64_t h1, h2;
    key_init(hdr | (1ull << 62), h1, h2);
    for (int q = lo; q < hi; q += 8) {
        uint64_t w = 0;
        const int n = min(8, hi - q);
        for (int i = 0; i < n; ++i) w |= (uint64_t)stored[q + i] << (8 * i);
        key_step(w, h1, h2);
    }
    key_finish(h1, h2);
    h[0] = h1; h[1] = h2;
}

// The two key hashes of a SCAR record from the RAW read (K4, verification).
// (weak: test hook - the segment is left out of the tags, so that different insertions collide and the exact
// verification has something to find)
__device__ __forceinline__ void key_hashes(const scar_record &rec, const uint8_t *raw, int rawlen, int platform,
                                           uint64_t &h1, uint64_t &h2, bool weak = false)
{
    const KeySegment k = key_segment(rec);
    key_init(k.hdr, h1, h2);
    if (weak) { key_finish(h1, h2); return; }
    bool acgt = true;
    for (int q = k.lo; q < k.hi && acgt; q += 32) {
        const int n = min(32, k.hi - q);
        uint64_t code = 0;
        for (int i = 0; i < n; ++i) {
            const uint8_t b = stored_byte(raw, rawlen, platform, q + i);
            acgt = acgt && scar_base_valid(b);
            code |= (uint64_t)scar_base_code(b) << (2 * i);
        }
        key_step(code, h1, h2);
    }
    if (acgt) { key_finish(h1, h2); return; }
    key_hashes_bytes(k, raw, rawlen, platform, h1, h2);
}

// ---- table ------------------------------------------------------------------------------------------
struct ScarTableRef {
    ScarTableSlot *slots;
    uint64_t slot_mask;      // capacity - 1
    uint32_t max_probe;
    int32_t *status;         // bit 1: table full
    uint32_t *claimed;       // indices of the claimed slots, in claim order ([capacity]; export / clear walk this list,
    unsigned long long *n_entries;   // not the whole table) and their number
};

__device__ __forceinline__ unsigned __int128 make_tag(uint64_t h1, uint64_t h2)
{
    return (unsigned __int128)h1 | ((unsigned __int128)h2 << 64);
}

// 16-byte load of a slot's tag that bypasses L1 (other SMs publish tags with L2 atomics)
__device__ __forceinline__ unsigned __int128 load_tag(const ScarTableSlot *s)
{
    unsigned long long lo, hi;
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(&s->tag));
    return (unsigned __int128)lo | ((unsigned __int128)hi << 64);
}

// Insert (or find) the key and fold n occurrences with smallest global index `first` into it.
// One slot visit = one plain 16-byte load; only an empty slot costs an atomic (one 128-bit CAS that
// publishes both hashes at once); count and first_idx are fire-and-forget reductions.
// Returns 1 when a new slot was claimed.
__device__ __forceinline__ int table_insert(const ScarTableRef &T, uint64_t h1, uint64_t h2, uint32_t n,
                                            uint64_t first, uint32_t sample, uint32_t &claimed_slot)
{
    const unsigned __int128 tag = make_tag(h1, h2);
    uint64_t slot = h1 & T.slot_mask;
    for (uint32_t probe = 0; probe <= T.max_probe; ++probe) {
        ScarTableSlot *s = T.slots + slot;
        unsigned __int128 cur = load_tag(s);
        int claimed = 0;
        // Both halves of a published tag are non-zero (key_finish), so a value with exactly one zero half can only
        // be a 16-byte load that was not single-copy atomic against a concurrent 128-bit CAS: read it again with an
        // atomic (compare == swap leaves the slot unchanged) instead of mistaking the slot for another key's.
        if (cur != 0 && ((uint64_t)cur == 0 || (uint64_t)(cur >> 64) == 0)) cur = atomicCAS(&s->tag, tag, tag);
        if (cur == 0) {
            cur = atomicCAS(&s->tag, (unsigned __int128)0, tag);
            claimed = cur == 0;
            if (claimed) cur = tag;
        }
        if (cur == tag) {
            atomicAdd(&s->count, n);
            atomicMax(&s->inv_first, ~(unsigned long long)first);
            if (claimed) { s->sample = sample; claimed_slot = (uint32_t)slot; }
            return claimed;
        }
        slot = (slot + 1) & T.slot_mask;
    }
    atomicOr(T.status, 2);
    return 0;
}

// Append the slots claimed by the lanes of a (converged, full) warp to the claimed-slot list: one atomic per warp.
__device__ __forceinline__ void table_note_claims(const ScarTableRef &T, bool claimed, uint32_t slot)
{
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, claimed);
    if (!mask) return;
    const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(T.n_entries, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (claimed) T.claimed[base + __popc(mask & ((1u << lane) - 1u))] = slot;
}

// ---- per-sample accumulators in shared memory (32-bit words) ------------------------------------------
// Six words of two 16-bit counters each (one atomic bumps two counters; a CTA flushes before any field can reach
// 2^16), HR_MORE, and the phase bins over PB = max phases + 2 values per side (none, phase 0 .. max-1, n/a): a joint
// [PB][PB] histogram of (phase_r1, phase_r2) (one atomic per read) or two [PB] marginals (two atomics, 2 PB words).
//   word A: ASSIGNED | PASSING << 16 | FILTERED << 32 | NO_JUNCTION << 48
//   word B: NO_CUT | N_INSERTION << 16 | SCAR << 32 | LEFT_DEL << 48
//   word C: RIGHT_DEL | INSERTION << 16 | MICROHOM << 32 | HR_FIRST << 48
#define K4_FIXED_WORDS 8
#define K4_EPOCH_ITERS 255          // 255 x 256 threads < 65536 reads per CTA between flushes

struct ScarAccRef {
    unsigned long long *counters;    // [n_samples][SCAR_N_COUNTERS]
    unsigned long long *phase_hist;  // [n_samples][2][SCAR_PHASE_BINS]
    unsigned long long *unassigned;
    int32_t n_samples;
    int32_t pb;                      // phase values per side: max phases + 2
    int32_t joint;                   // 1: joint [pb][pb] phase histogram, 0: two [pb] marginals
};

__host__ __device__ __forceinline__ int acc_words_per_sample(int pb, int joint) { return K4_FIXED_WORDS + (joint ? pb * pb : 2 * pb); }

// one read's contribution to the shared-memory accumulators of its sample
__device__ __forceinline__ void acc_add_smem(unsigned int *s_hist, int per_sample, int pb, int joint, const scar_record &rec)
{
    const uint32_t s = rec.sample;
    const bool scar = rec.status == SCAR_ST_SCAR;
    const int j1 = rec.phase_r1 == SCAR_PHASE_NA ? pb - 1 : rec.phase_r1 + 1;    // n/a is the last compact bin
    const int j2 = rec.phase_r2 == SCAR_PHASE_NA ? pb - 1 : rec.phase_r2 + 1;
    unsigned long long a = 1ull, b = 0ull, c = 0ull;
    if (rec.status >= SCAR_ST_NO_JUNCTION) a |= 1ull << 16;
    if (rec.status == SCAR_ST_FILTERED) a |= 1ull << 32;
    if (rec.status == SCAR_ST_NO_JUNCTION) a |= 1ull << 48;
    if (rec.status == SCAR_ST_NO_CUT) b |= 1ull;
    if (rec.status == SCAR_ST_N_INSERTION) b |= 1ull << 16;
    if (scar) {
        b |= 1ull << 32;
        if (rec.flags & SCAR_FL_LEFT_DEL) b |= 1ull << 48;
        if (rec.flags & SCAR_FL_RIGHT_DEL) c |= 1ull;
        if (rec.flags & SCAR_FL_INSERTION) c |= 1ull << 16;
        if (rec.flags & SCAR_FL_MICROHOM) c |= 1ull << 32;
    }
    if (rec.hr_n) c |= 1ull << 48;
    // 32-bit shared atomics (two 16-bit counters each); 64-bit shared atomics are a CAS loop
    unsigned int *h = s_hist + s * per_sample;
    atomicAdd(h, (unsigned int)a);
    if (a >> 32) atomicAdd(h + 1, (unsigned int)(a >> 32));
    if ((unsigned int)b) atomicAdd(h + 2, (unsigned int)b);
    if (b >> 32) atomicAdd(h + 3, (unsigned int)(b >> 32));
    if ((unsigned int)c) atomicAdd(h + 4, (unsigned int)c);
    if (c >> 32) atomicAdd(h + 5, (unsigned int)(c >> 32));
    if (rec.hr_n > 1) atomicAdd(h + 6, (unsigned)(rec.hr_n - 1));
    if (joint) atomicAdd(h + K4_FIXED_WORDS + j1 * pb + j2, 1u);
    else {
        atomicAdd(h + K4_FIXED_WORDS + j1, 1u);
        atomicAdd(h + K4_FIXED_WORDS + pb + j2, 1u);
    }
}

// the same straight into the global accumulators (more samples than shared memory holds)
__device__ __forceinline__ void acc_add_global(const ScarAccRef &A, const scar_record &rec)
{
    const uint32_t s = rec.sample;
    const int b1 = rec.phase_r1 == SCAR_PHASE_NA ? SCAR_PHASE_BINS - 1 : rec.phase_r1 + 1;
    const int b2 = rec.phase_r2 == SCAR_PHASE_NA ? SCAR_PHASE_BINS - 1 : rec.phase_r2 + 1;
    uint32_t inc = 1u << SCAR_CT_ASSIGNED;
    if (rec.status >= SCAR_ST_NO_JUNCTION) inc |= 1u << SCAR_CT_PASSING;
    if (rec.status == SCAR_ST_FILTERED) inc |= 1u << SCAR_CT_FILTERED;
    if (rec.status == SCAR_ST_NO_JUNCTION) inc |= 1u << SCAR_CT_NO_JUNCTION;
    if (rec.status == SCAR_ST_NO_CUT) inc |= 1u << SCAR_CT_NO_CUT;
    if (rec.status == SCAR_ST_N_INSERTION) inc |= 1u << SCAR_CT_N_INSERTION;
    if (rec.status == SCAR_ST_SCAR) {
        inc |= 1u << SCAR_CT_SCAR;
        if (rec.flags & SCAR_FL_LEFT_DEL) inc |= 1u << SCAR_CT_LEFT_DEL;
        if (rec.flags & SCAR_FL_RIGHT_DEL) inc |= 1u << SCAR_CT_RIGHT_DEL;
        if (rec.flags & SCAR_FL_INSERTION) inc |= 1u << SCAR_CT_INSERTION;
        if (rec.flags & SCAR_FL_MICROHOM) inc |= 1u << SCAR_CT_MICROHOM;
    }
    if (rec.hr_n) inc |= 1u << SCAR_CT_HR_FIRST;
    unsigned long long *cnt = A.counters + (size_t)s * SCAR_N_COUNTERS;
    for (uint32_t m = inc; m; m &= m - 1) atomicAdd(cnt + (__ffs(m) - 1), 1ull);
    if (rec.hr_n > 1) atomicAdd(cnt + SCAR_CT_HR_MORE, (unsigned long long)(rec.hr_n - 1));
    unsigned long long *ph = A.phase_hist + (size_t)s * 2 * SCAR_PHASE_BINS;
    atomicAdd(ph + b1, 1ull);
    atomicAdd(ph + SCAR_PHASE_BINS + b2, 1ull);
}

// flush the CTA's shared-memory accumulators into the global ones (all threads of the CTA; no barrier inside)
__device__ __forceinline__ void acc_flush(const ScarAccRef &A, unsigned int *s_hist, int per_sample)
{
    const int pb = A.pb;
    // counter id of 16-bit field f (0..11 in the order of the packed words A, B, C), 4 bits each
    constexpr unsigned long long FIELD_IDS =
        ((unsigned long long)SCAR_CT_ASSIGNED << 0) | ((unsigned long long)SCAR_CT_PASSING << 4) |
        ((unsigned long long)SCAR_CT_FILTERED << 8) | ((unsigned long long)SCAR_CT_NO_JUNCTION << 12) |
        ((unsigned long long)SCAR_CT_NO_CUT << 16) | ((unsigned long long)SCAR_CT_N_INSERTION << 20) |
        ((unsigned long long)SCAR_CT_SCAR << 24) | ((unsigned long long)SCAR_CT_LEFT_DEL << 28) |
        ((unsigned long long)SCAR_CT_RIGHT_DEL << 32) | ((unsigned long long)SCAR_CT_INSERTION << 36) |
        ((unsigned long long)SCAR_CT_MICROHOM << 40) | ((unsigned long long)SCAR_CT_HR_FIRST << 44);
    for (int i = threadIdx.x; i < A.n_samples * per_sample; i += blockDim.x) {
        const unsigned int v = s_hist[i];
        if (!v) continue;
        s_hist[i] = 0;
        const int s = i / per_sample, c = i - s * per_sample;
        unsigned long long *cnt = A.counters + (size_t)s * SCAR_N_COUNTERS;
        unsigned long long *ph = A.phase_hist + (size_t)s * 2 * SCAR_PHASE_BINS;
        if (c < 6) {                                   // half of a packed 64-bit word: two 16-bit fields
            const int f0 = (int)((FIELD_IDS >> (8 * c)) & 15ull), f1 = (int)((FIELD_IDS >> (8 * c + 4)) & 15ull);
            if (v & 0xFFFFu) atomicAdd(cnt + f0, (unsigned long long)(v & 0xFFFFu));
            if (v >> 16) atomicAdd(cnt + f1, (unsigned long long)(v >> 16));
        } else if (c == 6) {
            atomicAdd(cnt + SCAR_CT_HR_MORE, (unsigned long long)v);
        } else if (c >= K4_FIXED_WORDS) {
            const int b = c - K4_FIXED_WORDS;          // compact bin -> SCAR_PHASE_BINS layout (n/a is the last of both)
            if (A.joint) {
                const int b1 = b / pb, b2 = b - b1 * pb;
                atomicAdd(ph + (b1 == pb - 1 ? SCAR_PHASE_BINS - 1 : b1), (unsigned long long)v);
                atomicAdd(ph + SCAR_PHASE_BINS + (b2 == pb - 1 ? SCAR_PHASE_BINS - 1 : b2), (unsigned long long)v);
            } else {
                const int side = b >= pb, k = b - side * pb;
                atomicAdd(ph + side * SCAR_PHASE_BINS + (k == pb - 1 ? SCAR_PHASE_BINS - 1 : k), (unsigned long long)v);
            }
        }
    }
}

// The same without a barrier, for accumulators that other warps keep adding to: every word is emptied with an atomic
// exchange, so an increment lands either in this flush or in a later one.  `t` / `nt`: index and number of the
// calling threads.
__device__ __forceinline__ void acc_flush_exch(const ScarAccRef &A, unsigned int *s_hist, int per_sample, int t, int nt)
{
    const int pb = A.pb;
    constexpr unsigned long long FIELD_IDS =
        ((unsigned long long)SCAR_CT_ASSIGNED << 0) | ((unsigned long long)SCAR_CT_PASSING << 4) |
        ((unsigned long long)SCAR_CT_FILTERED << 8) | ((unsigned long long)SCAR_CT_NO_JUNCTION << 12) |
        ((unsigned long long)SCAR_CT_NO_CUT << 16) | ((unsigned long long)SCAR_CT_N_INSERTION << 20) |
        ((unsigned long long)SCAR_CT_SCAR << 24) | ((unsigned long long)SCAR_CT_LEFT_DEL << 28) |
        ((unsigned long long)SCAR_CT_RIGHT_DEL << 32) | ((unsigned long long)SCAR_CT_INSERTION << 36) |
        ((unsigned long long)SCAR_CT_MICROHOM << 40) | ((unsigned long long)SCAR_CT_HR_FIRST << 44);
    for (int i = t; i < A.n_samples * per_sample; i += nt) {
        if (!s_hist[i]) continue;
        const unsigned int v = atomicExch(s_hist + i, 0u);
        if (!v) continue;
        const int s = i / per_sample, c = i - s * per_sample;
        unsigned long long *cnt = A.counters + (size_t)s * SCAR_N_COUNTERS;
        unsigned long long *ph = A.phase_hist + (size_t)s * 2 * SCAR_PHASE_BINS;
        if (c < 6) {
            const int f0 = (int)((FIELD_IDS >> (8 * c)) & 15ull), f1 = (int)((FIELD_IDS >> (8 * c + 4)) & 15ull);
            if (v & 0xFFFFu) atomicAdd(cnt + f0, (unsigned long long)(v & 0xFFFFu));
            if (v >> 16) atomicAdd(cnt + f1, (unsigned long long)(v >> 16));
        } else if (c == 6) {
            atomicAdd(cnt + SCAR_CT_HR_MORE, (unsigned long long)v);
        } else if (c >= K4_FIXED_WORDS) {
            const int b = c - K4_FIXED_WORDS;
            if (A.joint) {
                const int b1 = b / pb, b2 = b - b1 * pb;
                atomicAdd(ph + (b1 == pb - 1 ? SCAR_PHASE_BINS - 1 : b1), (unsigned long long)v);
                atomicAdd(ph + SCAR_PHASE_BINS + (b2 == pb - 1 ? SCAR_PHASE_BINS - 1 : b2), (unsigned long long)v);
            } else {
                const int side = b >= pb, k = b - side * pb;
                atomicAdd(ph + side * SCAR_PHASE_BINS + (k == pb - 1 ? SCAR_PHASE_BINS - 1 : k), (unsigned long long)v);
            }
        }
    }
}
