// scar_window.cu — K3: phasing + read filters + sliding-window junction search + junction call.
//
// Replaces, for a whole batch, the per-read body of ScarSearch.data_processing
// (reference scarmapper/INDEL_Processing.py:135-196) and SlidingWindow.sliding_window
// (scarmapper/SlidingWindow.pyx:20-171), plus the phasing scan of consensus_demultiplex
// (INDEL_Processing.py:948-975).  Not a translation: the reference compares every 10-nt read
// window with every target window.  Here a window is a 20-bit key probed in a bucketed perfect-hash
// table (two entries per 8-byte bucket, one probe, no chains); the tables, the packed loci and the
// per-target metadata are one blob staged in shared memory by one TMA bulk copy per CTA.
//
// A CTA works on tiles of 256 reads.  The packed batch is cell-major ([cell][read]), so the tile's cells
// are W rows of 2 KB; for short reads they are copied into shared memory up front (cp.async).  Per tile:
//   phase 1  one thread per read: filters, phasing, the first anchor step of both junction searches;
//   phase 2  the searches that are still open are items in a shared-memory queue and are worked off in
//            rounds by dense warps (batch step + anchor step, re-queued while open);
//   phase 3  one thread per read: cut-window check, HR donor scan, N-in-insertion test, junction call,
//            one 16-byte record store.
// Two search strategies, both bit-exact: anchor-and-extend for loci whose 10-mers are all distinct
// (position table + 32-base packed compares), and an exhaustive cell scan for any locus (16 probes per
// cell issued back to back, first hit in scan order kept with a max / min of (position << 12 | i)).
//
//   left  search: positions p = len-20 .. 16 descending, first hit -> clj = p+10, tlj = cut - i
//   right search: positions p = 10 .. len-26 ascending,  first hit -> crj = p,    trj = cut + i
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "scar_table_core.cuh"

#ifndef SCAR_ANCHOR_BATCH
#define SCAR_ANCHOR_BATCH 14
#endif
#ifndef SCAR_K3_MIN_CTAS
#define SCAR_K3_MIN_CTAS 4
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// One TMA bulk copy (cp.async.bulk, SASS UBLKCP) of a blob into shared memory, completion signalled
// on an mbarrier; every thread of the CTA waits for it.
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

// Table access (layouts: ScarTableDesc in scar_internal.cuh).  In shared memory the table is addressed with 32-bit
// shared-window addresses and plain ld.shared (no generic-address conversion in the loop); otherwise it is read
// through the read-only global path.  probe(t), t = key << 12, returns the payload (<= 4094) on a hit and a value
// >= 4095 otherwise.
template <bool SMEM, bool COMPACT> struct TableRef;
template <> struct TableRef<true, false> {
    static constexpr bool compact = false;
    uint32_t base, mult, n;
    __device__ __forceinline__ TableRef(const void *smem_blob, const void *, const ScarTableDesc &d)
        : base(smem_u32(smem_blob) + d.off), mult(d.mult), n(d.n) {}
    __device__ __forceinline__ uint2 fetch(uint32_t t) const
    {
        uint2 v;
        // volatile: the loads of a batch stay in program order, ahead of their consumers
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(__umulhi(t * mult, n) * 8u + base));
        return v;
    }
};
template <> struct TableRef<false, false> {
    static constexpr bool compact = false;
    const uint2 *base; uint32_t mult, n;
    __device__ __forceinline__ TableRef(const void *, const void *gmem_blob, const ScarTableDesc &d)
        : base(reinterpret_cast<const uint2 *>(reinterpret_cast<const unsigned char *>(gmem_blob) + d.off)), mult(d.mult), n(d.n) {}
    __device__ __forceinline__ uint2 fetch(uint32_t t) const { return __ldg(base + __umulhi(t * mult, n)); }
};
template <> struct TableRef<true, true> {
    static constexpr bool compact = true;
    uint32_t ent, disp, mult, n, nb, unit;
    __device__ __forceinline__ TableRef(const void *smem_blob, const void *, const ScarTableDesc &d)
        : ent(smem_u32(smem_blob) + d.off), disp(smem_u32(smem_blob) + d.doff), mult(d.mult), n(d.n), nb(d.nb), unit(d.unit) {}
    __device__ __forceinline__ uint32_t fetch_disp(uint32_t h) const
    {
        uint32_t dv;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(dv) : "r"(scar_chd_bucket(h, nb) * 2u + disp));
        return dv;
    }
    __device__ __forceinline__ uint32_t fetch_entry(uint32_t h, uint32_t dv) const
    {
        uint32_t e;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(scar_chd_slot(h, dv, unit, n) * 4u + ent));
        return e;
    }
};
template <> struct TableRef<false, true> {
    static constexpr bool compact = true;
    const uint32_t *ent; const uint16_t *disp; uint32_t mult, n, nb, unit;
    __device__ __forceinline__ TableRef(const void *, const void *gmem_blob, const ScarTableDesc &d)
        : ent(reinterpret_cast<const uint32_t *>(reinterpret_cast<const unsigned char *>(gmem_blob) + d.off)),
          disp(reinterpret_cast<const uint16_t *>(reinterpret_cast<const unsigned char *>(gmem_blob) + d.doff)),
          mult(d.mult), n(d.n), nb(d.nb), unit(d.unit) {}
    __device__ __forceinline__ uint32_t fetch_disp(uint32_t h) const { return __ldg(disp + scar_chd_bucket(h, nb)); }
    __device__ __forceinline__ uint32_t fetch_entry(uint32_t h, uint32_t dv) const { return __ldg(ent + scar_chd_slot(h, dv, unit, n)); }
};

// one probe: the payload (<= 4094) on a hit and a value >= 4095 otherwise; t = key << 12
template <class TAB>
__device__ __forceinline__ uint32_t probe_one(const TAB &tab, uint32_t t)
{
    if constexpr (TAB::compact) {
        const uint32_t h = t * tab.mult;
        return tab.fetch_entry(h, tab.fetch_disp(h)) ^ t;
    } else {
        const uint2 v = tab.fetch(t);
        return min(v.x ^ t, v.y ^ t);
    }
}

// N consecutive windows starting at the low end of (wlo, whi).  t[o] holds the window's 20 key bits in its TOP bits
// (raw << 12); the loads of all N probes are issued back to back (N independent shared-memory loads in flight per
// thread) before any result is formed; m[o] = payload (< 4095) on a hit and a value >= 4095 otherwise.
template <class TAB, int N>
__device__ __forceinline__ void probe_n(const TAB &tab, uint32_t wlo, uint32_t whi, uint32_t (&m)[N])
{
    uint32_t t[N];
    if constexpr (TAB::compact) {
        uint32_t h[N], dv[N];
#pragma unroll
        for (int o = 0; o < N; ++o) {
            t[o] = __funnelshift_r(wlo, whi, 2 * o) << 12;
            h[o] = t[o] * tab.mult;
            dv[o] = tab.fetch_disp(h[o]);
        }
#pragma unroll
        for (int o = 0; o < N; ++o) m[o] = tab.fetch_entry(h[o], dv[o]);
#pragma unroll
        for (int o = 0; o < N; ++o) m[o] ^= t[o];
    } else {
        uint2 e[N];
#pragma unroll
        for (int o = 0; o < N; ++o) {
            t[o] = __funnelshift_r(wlo, whi, 2 * o) << 12;
            e[o] = tab.fetch(t[o]);
        }
#pragma unroll
        for (int o = 0; o < N; ++o) m[o] = min(e[o].x ^ t[o], e[o].y ^ t[o]);
    }
}

// the 16 windows of one cell; returns min(m): "any window of this cell is in the table" costs one compare per
// cell, and the scan-order resolution runs only then
template <class TAB>
__device__ __forceinline__ uint32_t probe16(const TAB &tab, uint32_t wlo, uint32_t whi, uint32_t (&m)[16])
{
    probe_n<TAB, 16>(tab, wlo, whi, m);
    const uint32_t a = min(min(m[0], m[1]), min(m[2], m[3])), b = min(min(m[4], m[5]), min(m[6], m[7]));
    const uint32_t c = min(min(m[8], m[9]), min(m[10], m[11])), d = min(min(m[12], m[13]), min(m[14], m[15]));
    return min(min(a, b), min(c, d));
}

// bit o set iff bases o .. o+9 of the 32-base window are all ACGT
__device__ __forceinline__ uint32_t runs_of_ten(uint32_t vwin)
{
    const uint32_t a = vwin & (vwin >> 1);      // runs of 2
    const uint32_t b = a & (a >> 2);            // runs of 4
    const uint32_t c = b & (b >> 4);            // runs of 8
    return c & (a >> 8);                        // runs of 10
}

// ---- anchor-and-extend (loci whose 10-mers are all distinct) -----------------------------------------
// A probe of the position table tells where in the locus a read window lies.  When it lies on the far
// side of the cut, the flank the scan is walking through is compared with the locus 16 bases per XOR
// (2-bit packed words); every window inside the verified stretch is a locus window at a known
// position, and because each 10-mer occurs once in the locus, whether it is a left / right window is
// decided by that position alone — so the scan jumps over the stretch instead of probing each window.
// Positions around a mismatch or the junction are still probed one by one: results are identical to
// the exhaustive scan.
#ifndef K3_TILE
#define K3_TILE 256
#endif
#ifndef K3_STAGE_MAX_CELLS
#define K3_STAGE_MAX_CELLS 12          // reads up to 192 nt: 24 KB of cells per tile
#endif

// The cells of one read.  ST: the tile's cells were copied to shared memory (cell j of local read rl at
// sc[j * K3_TILE]); otherwise they are read through the read-only global path (cell-major, stride S).
template <bool ST> struct ReadCursorT {
    const uint64_t *col; int64_t S; const uint64_t *sc; int Wr;
    __device__ __forceinline__ uint64_t cell(int j) const { return ST ? sc[j * K3_TILE] : __ldg(col + (int64_t)j * S); }
    __device__ __forceinline__ uint64_t cell_or_zero(int j) const { return j < Wr ? cell(j) : 0ull; }
    // 16 bases starting at base b >= 0: 2-bit codes and ACGT mask
    __device__ __forceinline__ void word16(int b, uint32_t &code, uint32_t &valid) const
    {
        const int j = b >> 4, o = b & 15;
        const uint64_t c0 = cell_or_zero(j), c1 = cell_or_zero(j + 1);
        code = __funnelshift_r((uint32_t)c0, (uint32_t)c1, 2 * o);
        valid = __funnelshift_r((uint32_t)(c0 >> 32) << 16, (uint32_t)(c1 >> 32), 16 + o) & 0xFFFFu;
    }
};

// 16 bases of the packed locus starting at base b (b >= -16; the blob pads one word in front)
template <bool SMEM>
__device__ __forceinline__ uint32_t locus16(const uint32_t *tw, int b)
{
    const int bb = b + 16, w = bb >> 4, o = bb & 15;
    uint32_t lo, hi;
    if (SMEM) {
        const uint32_t a = smem_u32(tw) + (uint32_t)w * 4u;
        asm("ld.shared.u32 %0, [%1];" : "=r"(lo) : "r"(a));
        asm("ld.shared.u32 %0, [%1];" : "=r"(hi) : "r"(a + 4u));
    } else { lo = __ldg(tw + w); hi = __ldg(tw + w + 1); }
    return __funnelshift_r(lo, hi, 2 * o);
}

// mismatch bits (bit 2k set: base k differs or the read base is not ACGT)
__device__ __forceinline__ uint32_t mismatch16(uint32_t rcode, uint32_t rvalid, uint32_t tcode)
{
    const uint32_t x = rcode ^ tcode;
    uint32_t mm = (x | (x >> 1)) & 0x55555555u;
    if (rvalid != 0xFFFFu) {                      // rare: a non-ACGT read base counts as a mismatch
        uint32_t inv = ~rvalid & 0xFFFFu;
        inv = (inv | (inv << 8)) & 0x00FF00FFu; inv = (inv | (inv << 4)) & 0x0F0F0F0Fu;
        inv = (inv | (inv << 2)) & 0x33333333u; inv = (inv | (inv << 1)) & 0x55555555u;
        mm |= inv;
    }
    return mm;
}

// 32 bases starting at base b >= 0 of the read (2-bit codes, ACGT mask) and of the locus
struct Word32 { uint64_t code; uint32_t valid; };
template <class RC>
__device__ __forceinline__ Word32 read32(const RC &rc, int b)
{
    const int j = b >> 4, o = b & 15;
    const uint64_t c0 = rc.cell_or_zero(j), c1 = rc.cell_or_zero(j + 1), c2 = rc.cell_or_zero(j + 2);
    Word32 w;
    const uint32_t lo = __funnelshift_r((uint32_t)c0, (uint32_t)c1, 2 * o);
    const uint32_t hi = __funnelshift_r((uint32_t)c1, (uint32_t)c2, 2 * o);
    w.code = ((uint64_t)hi << 32) | lo;
    const uint64_t v48 = ((c0 >> 32) & 0xFFFFull) | (((c1 >> 32) & 0xFFFFull) << 16) | (((c2 >> 32) & 0xFFFFull) << 32);
    w.valid = (uint32_t)(v48 >> o);
    return w;
}

template <bool SMEM>
__device__ __forceinline__ uint64_t locus32(const uint32_t *tw, int b)     // b >= -16
{
    const int bb = b + 16, w = bb >> 4, o = bb & 15;
    uint32_t a0, a1, a2;
    if (SMEM) {
        const uint32_t a = smem_u32(tw) + (uint32_t)w * 4u;
        asm("ld.shared.u32 %0, [%1];" : "=r"(a0) : "r"(a));
        asm("ld.shared.u32 %0, [%1];" : "=r"(a1) : "r"(a + 4u));
        asm("ld.shared.u32 %0, [%1];" : "=r"(a2) : "r"(a + 8u));
    } else { a0 = __ldg(tw + w); a1 = __ldg(tw + w + 1); a2 = __ldg(tw + w + 2); }
    return ((uint64_t)__funnelshift_r(a1, a2, 2 * o) << 32) | __funnelshift_r(a0, a1, 2 * o);
}

// mismatch bits of 32 bases (bit 2k set: base k differs or the read base is not ACGT)
__device__ __forceinline__ uint64_t mismatch32(const Word32 &r, uint64_t tcode)
{
    const uint64_t x = r.code ^ tcode;
    uint64_t mm = (x | (x >> 1)) & 0x5555555555555555ull;
    if (r.valid != 0xFFFFFFFFu) {                 // rare: a non-ACGT read base counts as a mismatch
        uint64_t inv = (uint64_t)(~r.valid);
        inv = (inv | (inv << 16)) & 0x0000FFFF0000FFFFull; inv = (inv | (inv << 8)) & 0x00FF00FF00FF00FFull;
        inv = (inv | (inv << 4)) & 0x0F0F0F0F0F0F0F0Full; inv = (inv | (inv << 2)) & 0x3333333333333333ull;
        inv = (inv | (inv << 1)) & 0x5555555555555555ull;
        mm |= inv;
    }
    return mm;
}

// number of bases, going LEFT from read base q-1 / locus base tq-1, that match; at most n.
// The window may start before base 0 of the read on the last step (q-k >= 16 only guarantees 16 bases), so
// a step is clamped to what exists.
template <bool SMEM, class RC>
__device__ __forceinline__ int match_left(const RC &rc, const uint32_t *tw, int q, int tq, int n)
{
    int k = 0;
    while (k < n) {
        const int e = q - k;                                      // bases e-w .. e-1, w = min(32, e)
        const int w = min(32, min(e, tq - k + 16));
        const Word32 r = read32(rc, e - w);
        uint64_t mm = mismatch32(r, locus32<SMEM>(tw, tq - k - w));
        mm <<= 2 * (32 - w);                                      // align the top base to bit 62
        const int run = mm ? (__clzll((long long)mm) >> 1) : 32;  // matches counted from the top base down
        const int c = min(w, n - k);
        if (run < c) return k + run;
        k += c;
    }
    return n;
}

// number of bases, going RIGHT from read base q / locus base tq, that match; at most n
template <bool SMEM, class RC>
__device__ __forceinline__ int match_right(const RC &rc, const uint32_t *tw, int q, int tq, int n)
{
    int k = 0;
    while (k < n) {
        const Word32 r = read32(rc, q + k);
        const uint64_t mm = mismatch32(r, locus32<SMEM>(tw, tq + k));
        const int run = mm ? ((__ffsll((long long)mm) - 1) >> 1) : 32;
        const int c = min(32, n - k);
        if (run < c) return k + run;
        k += c;
    }
    return n;
}

// position in the locus of the read window at p, or 0xFFFFFFFF (not a locus 10-mer / not all ACGT)
template <class TAB, class RC>
__device__ __forceinline__ uint32_t locus_position(const RC &rc, const TAB &ptab, int p)
{
    uint32_t code, valid;
    rc.word16(p, code, valid);
    if ((valid & 0x3FFu) != 0x3FFu) return 0xFFFFFFFFu;
    const uint32_t m = probe_one(ptab, code << 12);
    return m < 4095u ? m : 0xFFFFFFFFu;
}

__device__ __forceinline__ void store_record(scar_record *dst, const scar_record &rec)
{
    *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(&rec);
}

// ---- anchor-and-extend steps.  A junction search alternates two steps until it is resolved:
//   anchor: one probe of the position table at p; when the window lies on the far side of the cut the flank
//           is verified against the locus 32 bases per XOR and the scan jumps over it;
//   batch : the next SCAR_ANCHOR_BATCH positions probed back to back (around a mismatch or the junction).
// Each returns true while the search is still open (p updated) and false when it is finished (clj/tlj or
// crj/trj set, or left at "not found").
struct LocusRef {                  // what a step needs to know about the read's target
    const uint32_t *tw;
    int cut, L;
};

template <bool SMEM, class TAB, class RC>
__device__ __forceinline__ bool left_anchor(const RC &rc, const TAB &ptab, const LocusRef &T, int &p,
                                            int &clj, int &tlj)
{
    // left junction (SlidingWindow.pyx:49-69): the largest p in 16 .. len-20 whose window is a left window,
    // i.e. lies at locus position 6 <= t <= cut-10 (window i = cut-10-t)
    const uint32_t t = locus_position(rc, ptab, p);
    if (t != 0xFFFFFFFFu && (int)t > T.cut - 10) {
        const int need = (int)t - (T.cut - 10), limit = min(need, p - 16);
        const int k = match_left<SMEM>(rc, T.tw, p, (int)t, limit);
        if (k == limit) {                              // the whole stretch is locus sequence
            if (limit == need) { clj = p - need + 10; tlj = T.cut; }
            return false;                              // else: ran out of read positions
        }
        p -= k + 1;                                    // windows p-1 .. p-k are right of the cut
        return true;
    }
    if (t != 0xFFFFFFFFu && t >= 6u) { clj = p + 10; tlj = (int)t + 10; return false; }
    return --p >= 16;
}

template <bool SMEM, class TAB, class RC>
__device__ __forceinline__ bool left_batch(const RC &rc, const TAB &ptab, const LocusRef &T, int &p,
                                           int &clj, int &tlj)
{
    // batch width: a clean deletion's junction window is the 10th position past the break, 1-3 nt
    // insertions push it to the 11th-13th; wider batches cost every read, narrower ones cost rounds
    constexpr int NB = SCAR_ANCHOR_BATCH;
    constexpr uint32_t NBMASK = (1u << NB) - 1u;
    const int b = p - (NB - 1);                        // positions p, p-1, .. b; b >= 0 because p >= 16
    uint32_t wlo, whi, vlo, vhi, m[NB];
    rc.word16(b + 16, whi, vhi);
    rc.word16(b, wlo, vlo);
    probe_n<TAB, NB>(ptab, wlo, whi, m);
    // left window iff (t - 6) <= cut-16: one bit per position, then the first in scan order
    uint32_t hm = 0;
#pragma unroll
    for (int o = 0; o < NB; ++o) hm |= ((m[o] - 6u) <= (uint32_t)(T.cut - 16)) ? (1u << o) : 0u;
    if (hm) hm &= runs_of_ten(vlo | (vhi << 16)) & (b >= 16 ? NBMASK : ((NBMASK << (16 - b)) & NBMASK));
    if (hm) {
        const int o = 31 - __clz(hm);                  // highest position = first going down
        uint32_t t = 0;
#pragma unroll
        for (int q = 0; q < NB; ++q) t = (q == o) ? m[q] : t;
        clj = b + o + 10; tlj = (int)t + 10;
        return false;
    }
    p = b - 1;
    return p >= 16;
}

template <bool SMEM, class TAB, class RC>
__device__ __forceinline__ bool right_anchor(const RC &rc, const TAB &ptab, const LocusRef &T, int plast,
                                             int &p, int &crj, int &trj)
{
    // right junction (SlidingWindow.pyx:79-93): the smallest p in 10 .. len-26 whose window is a right window,
    // i.e. lies at locus position cut <= t <= L-16 (window i = t-cut)
    const uint32_t t = locus_position(rc, ptab, p);
    if (t != 0xFFFFFFFFu && (int)t < T.cut) {
        const int need = T.cut - (int)t, limit = min(need, plast - p);
        const int k = match_right<SMEM>(rc, T.tw, p + 10, (int)t + 10, limit);
        if (k == limit) {
            if (limit == need) { crj = p + need; trj = T.cut; }
            return false;
        }
        p += k + 1;                                    // windows p+1 .. p+k are left of the cut
        return true;
    }
    if (t != 0xFFFFFFFFu && (int)t <= T.L - 16) { crj = p; trj = (int)t; return false; }
    return ++p <= plast;
}

template <bool SMEM, class TAB, class RC>
__device__ __forceinline__ bool right_batch(const RC &rc, const TAB &ptab, const LocusRef &T, int plast,
                                            int &p, int &crj, int &trj)
{
    constexpr int NB = SCAR_ANCHOR_BATCH;
    constexpr uint32_t NBMASK = (1u << NB) - 1u;
    uint32_t wlo, whi, vlo, vhi, m[NB];
    rc.word16(p, wlo, vlo);                            // positions p .. p+NB-1
    rc.word16(p + 16, whi, vhi);
    probe_n<TAB, NB>(ptab, wlo, whi, m);
    const uint32_t span = (uint32_t)(T.L - 16 - T.cut);   // right window iff (t - cut) <= span
    uint32_t hm = 0;
#pragma unroll
    for (int o = 0; o < NB; ++o) hm |= ((m[o] - (uint32_t)T.cut) <= span) ? (1u << o) : 0u;
    if (hm) {
        const int top = plast - p;                     // >= 0
        hm &= runs_of_ten(vlo | (vhi << 16)) & (top >= NB - 1 ? NBMASK : ((2u << top) - 1u));
    }
    if (hm) {
        const int o = __ffs(hm) - 1;                   // lowest position = first going up
        uint32_t t = 0;
#pragma unroll
        for (int q = 0; q < NB; ++q) t = (q == o) ? m[q] : t;
        crj = p + o; trj = (int)t;
        return false;
    }
    p += NB;
    return p <= plast;
}

// append the lanes with `pending` to a CTA queue: one shared-memory atomic per warp
__device__ __forceinline__ void queue_push(bool pending, uint32_t item, uint32_t *queue, int dir, uint32_t *counter)
{
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, pending);
    if (!mask) return;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == __ffs(mask) - 1) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(mask) - 1);
    if (pending) queue[dir * (int)(base + __popc(mask & ((1u << lane) - 1u)))] = item;
}

// ---- fused mode: raw bytes are staged in shared memory and packed there (what K1 does into HBM) -----------------
// Every WARP runs its own small pipeline over the 32 reads it owns in a tile: chunks of FUSED_CHUNK reads are brought
// into warp-private slots by TMA bulk copies (lane 0 issues, a per-slot mbarrier signals the bytes), FUSED_LANES
// lanes pack one read (a run of consecutive cells each), and the slot is refilled with the chunk FUSED_SLOTS ahead -
// the first chunk(s) of the NEXT tile while this tile is being searched.  No CTA barrier is involved: the cells a
// warp writes are read by other warps only between the barriers of the search rounds.
// LANES (template parameter): lanes per read = chunks per warp per tile; a chunk holds 32 / LANES reads.
#ifndef FUSED_SLOTS
#define FUSED_SLOTS 1
#endif

struct ChunkDesc { unsigned long long lo16; uint32_t span; bool staged; };

// raw byte span of reads [r0, r0 + chunk_reads) (uniform over the warp)
__device__ __forceinline__ ChunkDesc chunk_desc(const scar_ragged &seq, int64_t r0, int chunk_reads, int64_t n_reads, uint32_t slot_bytes)
{
    ChunkDesc d; d.lo16 = 0; d.span = 0; d.staged = false;
    if (r0 >= n_reads) return d;
    const int64_t r1 = min(r0 + chunk_reads, n_reads) - 1;
    unsigned long long lo, hi;
    if (seq.offsets) {
        lo = (unsigned long long)(uintptr_t)seq.bytes + (unsigned long long)seq.offsets[r0];
        hi = (unsigned long long)(uintptr_t)seq.bytes + (unsigned long long)(seq.lengths ? seq.offsets[r1] + seq.lengths[r1] : seq.offsets[r1 + 1]);
    } else {
        // fixed stride: every read owns `stride` bytes of the buffer, so the span may end at the last read's stride
        // boundary whatever its length (no load of lengths[r1] in front of the copy)
        lo = (unsigned long long)(uintptr_t)seq.bytes + (unsigned long long)(r0 * seq.stride);
        hi = lo + (unsigned long long)((r1 - r0 + 1) * seq.stride);
    }
    d.lo16 = lo & ~15ull;
    if (hi > d.lo16) {
        const unsigned long long span = (hi - d.lo16 + 15ull) & ~15ull;
        d.staged = span + 32ull <= (unsigned long long)slot_bytes;      // 32 bytes of slack: loads past a read's end
        d.span = (uint32_t)span;
    }
    return d;
}

// ask L2 for a span that a later TMA copy will take (the chunk after the one in flight)
__device__ __forceinline__ void chunk_prefetch_l2(const ChunkDesc &d)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(d.lo16), "r"(d.span) : "memory");
}

__device__ __forceinline__ void chunk_issue(const ChunkDesc &d, uint32_t slot_addr, uint32_t bar_addr)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(d.span) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(slot_addr), "l"(d.lo16), "r"(d.span), "r"(bar_addr) : "memory");
}

__device__ __forceinline__ void chunk_wait(uint32_t bar_addr, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar_addr), "r"(parity) : "memory");
    }
}

// cell j of a read straight from global memory, one byte at a time (a chunk whose span does not fit its slot:
// scattered views, reads longer than the context was sized for); kept out of line - the hot path is pack_chunk's loop
static __device__ __noinline__ uint2 pack_cell_bytes(const uint8_t *sp, int slen, int j)
{
    uint32_t code = 0u, valid = 0u, nmask = 0u;
    for (int k = 0; k < 16 && 16 * j + k < slen; ++k) {
        const uint8_t b = sp[16 * j + k];
        if (scar_base_valid(b)) { valid |= 1u << k; code |= scar_base_code(b) << (2 * k); }
        if (b == 'N') nmask |= 1u << k;
    }
    return make_uint2(code, valid | (nmask << 16));
}

// Pack the reads of one chunk: lane -> read k = lane / LANES of the chunk, cells [q * cpl, (q + 1) * cpl) with
// q = lane % LANES and cpl = ceil(W / LANES).  `lr0` is the tile-local index of the chunk's first read;
// s_info[local read] receives stored length | N count << 16.
// byte masks of the first nb (< 16) bytes of a 16-byte cell, one uint4 per nb (last cell of a read)
__constant__ uint4 SCAR_KEEP_BYTES[16] = {
    {0x00000000u, 0u, 0u, 0u}, {0x000000FFu, 0u, 0u, 0u}, {0x0000FFFFu, 0u, 0u, 0u}, {0x00FFFFFFu, 0u, 0u, 0u},
    {0xFFFFFFFFu, 0x00000000u, 0u, 0u}, {0xFFFFFFFFu, 0x000000FFu, 0u, 0u}, {0xFFFFFFFFu, 0x0000FFFFu, 0u, 0u}, {0xFFFFFFFFu, 0x00FFFFFFu, 0u, 0u},
    {0xFFFFFFFFu, 0xFFFFFFFFu, 0x00000000u, 0u}, {0xFFFFFFFFu, 0xFFFFFFFFu, 0x000000FFu, 0u}, {0xFFFFFFFFu, 0xFFFFFFFFu, 0x0000FFFFu, 0u}, {0xFFFFFFFFu, 0xFFFFFFFFu, 0x00FFFFFFu, 0u},
    {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x00000000u}, {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x000000FFu}, {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x0000FFFFu}, {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x00FFFFFFu}};

// exact masks of a cell that holds a non-ACGT byte (rare; out of line so that the pack loop stays small):
// returns code | (ACGT mask | N mask << 16) << 32, `hi_in` = mask of the bases that exist
static __device__ __noinline__ uint2 pack_cell_masks(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3, uint32_t code, uint32_t hi_in)
{
    const uint32_t v[4] = {v0, v1, v2, v3};
    uint32_t valid = 0u, nmask = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        valid |= flags_to_nibble(zero_bytes(acgt_diff4(v[i]))) << (4 * i);
        nmask |= flags_to_nibble(zero_bytes(v[i] ^ 0x4E4E4E4Eu)) << (4 * i);
    }
    valid &= hi_in; nmask &= hi_in;
    // spread the ACGT mask to 2 bits per base to clear the codes of non-ACGT bases
    uint32_t m2 = valid;
    m2 = (m2 | (m2 << 8)) & 0x00FF00FFu; m2 = (m2 | (m2 << 4)) & 0x0F0F0F0Fu;
    m2 = (m2 | (m2 << 2)) & 0x33333333u; m2 = (m2 | (m2 << 1)) & 0x55555555u;
    return make_uint2(code & (m2 * 3u), valid | (nmask << 16));
}

template <int LANES>
__device__ __forceinline__ void pack_chunk(const K3Params &P, const ChunkDesc &d, uint32_t slot_addr, int64_t r0, int lr0,
                                           uint64_t *s_cells, uint32_t *s_info)
{
    const int lane = threadIdx.x & 31, k = lane / LANES, q = lane % LANES;
    const int64_t r = r0 + k;
    const bool have = r < P.n_reads;
    const uint8_t *rp = P.seq.bytes; int rawlen = 0;
    if (have) get_item(P.seq, r, rp, rawlen);
    int slen = rawlen, shift = 0;
    if (P.platform == SCAR_PLATFORM_TRUSEQ) { slen = rawlen > 12 ? rawlen - 12 : 0; shift = 6; }
    if (slen > 16 * P.W) { if (q == 0) atomicOr(P.status, 1); slen = 16 * P.W; }
    const uint8_t *sp = rp + shift;
    // in the staged span?  (views that are not in file order may fall outside: those are read from global memory)
    const unsigned long long a = (unsigned long long)(uintptr_t)sp;
    const bool inside = d.staged && a >= d.lo16 && (unsigned long long)(uintptr_t)rp + (unsigned long long)rawlen <= d.lo16 + d.span;
    const int cpl = (P.W + LANES - 1) / LANES;
    const int j0 = q * cpl, j1 = min(j0 + cpl, P.W);
    uint2 *out = reinterpret_cast<uint2 *>(s_cells + lr0 + k) + (size_t)j0 * K3_TILE;
    int n_count = 0;
    if (inside && have) {                                   // (a lane without a read takes the other branch: no store test per cell here)
        const uint32_t a0 = slot_addr + (uint32_t)(a - d.lo16) + 16u * (uint32_t)j0;     // shared address of cell j0's first byte
        const uint32_t sh = (a0 & 3u) * 8u;
        uint32_t a4 = a0 & ~3u, wp;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wp) : "r"(a4));
        int nb = slen - 16 * j0;                            // bases from this cell on
        // one cell: four words of the raw bytes -> 16 codes + masks.  PARTIAL: the read ends in (or before) this cell
        auto pack_cell = [&](auto partial) {
            uint32_t w1, w2, w3, w4;
            asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(a4));
            asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(a4));
            asm volatile("ld.shared.u32 %0, [%1+12];" : "=r"(w3) : "r"(a4));
            asm volatile("ld.shared.u32 %0, [%1+16];" : "=r"(w4) : "r"(a4));
            uint32_t v0 = __funnelshift_r(wp, w1, sh), v1 = __funnelshift_r(w1, w2, sh);
            uint32_t v2 = __funnelshift_r(w2, w3, sh), v3 = __funnelshift_r(w3, w4, sh);
            wp = w4;
            uint32_t hi = 0xFFFFu;                          // ACGT mask | N mask << 16
            if (decltype(partial)::value) {                 // bytes past the read's end count as 'A' (code 0)
                const int nbc = max(nb, 0);
                hi = (1u << nbc) - 1u;
                const uint4 keep = SCAR_KEEP_BYTES[nbc];
                v0 = (v0 & keep.x) | (0x41414141u & ~keep.x); v1 = (v1 & keep.y) | (0x41414141u & ~keep.y);
                v2 = (v2 & keep.z) | (0x41414141u & ~keep.z); v3 = (v3 & keep.w) | (0x41414141u & ~keep.w);
            }
            const uint32_t x = acgt_any16(v0, v1, v2, v3);
            uint2 cell;
            cell.x = __byte_perm(__byte_perm(codes_to_top_byte(v0), codes_to_top_byte(v1), 0x0073),
                                 __byte_perm(codes_to_top_byte(v2), codes_to_top_byte(v3), 0x0073), 0x5410);
            cell.y = hi;
            if (x != 0u) {                                  // rare: a non-ACGT byte in the cell
                cell = pack_cell_masks(v0, v1, v2, v3, cell.x, hi);
                n_count += __popc(cell.y >> 16);
            }
            *out = cell;
        };
#ifdef SCAR_PACK_ONE_LOOP
        constexpr bool split = false;
#elif defined(SCAR_PACK_SPLIT_LOOP)
        constexpr bool split = true;
#else
        // measured (profiles/r02/pack_variants_sweep.txt): with 2 lanes per read the split loops take 6.8 % off the whole
        // kernel on C2, with 4 lanes per read (reads over 160 nt) they cost 2.3 % on C4
        constexpr bool split = LANES == 2;
#endif
        if (split) {
            // the cells that lie wholly inside the read need no end-of-read test; then the cell the read ends in and
            // whatever this lane still owns behind it
            const int n_cells = j1 - j0, n_full = max(0, min(n_cells, nb >> 4));
            int j = 0;
#ifdef SCAR_PACK_UNROLL
            constexpr int pack_unroll = SCAR_PACK_UNROLL;
#pragma unroll pack_unroll
#endif
            for (; j < n_full; ++j, a4 += 16u, out += K3_TILE) pack_cell(std::false_type{});
            nb -= 16 * n_full;
#ifdef SCAR_PACK_PARTIAL_UNROLL1
#pragma unroll 1
#endif
            for (; j < n_cells; ++j, a4 += 16u, nb -= 16, out += K3_TILE) pack_cell(std::true_type{});
        } else {
            for (int j = j0; j < j1; ++j, a4 += 16u, nb -= 16, out += K3_TILE) {
                if (nb < 16) pack_cell(std::true_type{}); else pack_cell(std::false_type{});
            }
        }
    } else {
        for (int j = j0; j < j1; ++j, out += K3_TILE) {
            const uint2 c = have ? pack_cell_bytes(sp, slen, j) : make_uint2(0u, 0u);
            n_count += __popc(c.y >> 16);
            if (have) *out = c;
        }
    }
#pragma unroll
    for (int m = 1; m < LANES; m <<= 1) n_count += __shfl_xor_sync(0xFFFFFFFFu, n_count, m);
    if (q == 0) s_info[lr0 + k] = have ? ((uint32_t)slen | ((uint32_t)n_count << 16)) : 0u;
}

// MODE 0: the packed cells are read from global memory (cell-major)          } input = K1's packed batch
// MODE 1: the tile's packed cells are first copied into shared memory        }
// MODE 2: FUSED - input = raw read bytes: the tile's raw span is staged in shared memory by TMA bulk copies, packed
//         there (the cells overwrite the raw bytes), classified, and in phase 3 the per-sample accumulators and the
//         scar table are updated as well (what K1, K3 and K4 do in three passes over HBM).
// groups of 256 threads per CTA of the fused kernel (registers: 64 per thread with 4 groups, 128 with 2)
__host__ __device__ constexpr int fused_groups_max(int lanes) { return lanes > 2 ? 2 : 4; }
template <bool SMEM, bool HR, int MODE, int LANES, bool COMPACT>
__global__ void __launch_bounds__(MODE == 2 ? fused_groups_max(LANES) * K3_TILE : K3_TILE, MODE == 2 ? 1 : SCAR_K3_MIN_CTAS)
scar_window_kernel(const K3Params P)
{
    constexpr int FUSED_CHUNK = 32 / LANES, FUSED_CHUNKS = LANES;      // reads per chunk, chunks per warp per tile
    // FUSED: a CTA is up to GMAX independent GROUPS of 256 threads (one tile each at a time, named barriers) that share
    // the static blob and the accumulators in shared memory - the occupancy of GMAX small CTAs without GMAX copies of
    // the tables.  tx = thread index inside the group; the other modes have one group.
    constexpr int GMAX = MODE == 2 ? fused_groups_max(LANES) : 1;
    const int group = GMAX > 1 ? (int)(threadIdx.x / K3_TILE) : 0, tx = GMAX > 1 ? (int)(threadIdx.x % K3_TILE) : (int)threadIdx.x;
    const int n_groups = GMAX > 1 ? (int)(blockDim.x / K3_TILE) : 1;
    auto group_sync = [&]() {
        if (GMAX > 1) asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(K3_TILE) : "memory");
        else __syncthreads();
    };
    constexpr bool ST = MODE >= 1;      // the tile's cells live in shared memory
    constexpr bool FUSED = MODE == 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [mbarriers: blob + (fused) one per warp and ring slot][static blob: tables | target metas | phases |
    // N-limit table][tile buffer: cells (+ fused: raw rings)][fused: accumulators]
    constexpr uint32_t HDR = FUSED ? 16u + GMAX * (K3_TILE / 32) * FUSED_SLOTS * 8u : 16u;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    const unsigned char *blob = reinterpret_cast<const unsigned char *>(P.blob);
    if (SMEM) {
        stage_tables_tma(smem_raw + HDR, P.blob, P.blob_bytes, bar);
        blob = smem_raw + HDR;
    }
    const ScarTargetMeta *metas = reinterpret_cast<const ScarTargetMeta *>(blob + P.meta_off);
    const ScarPhase *phases = reinterpret_cast<const ScarPhase *>(blob + P.phase_off);
    const uint16_t *nmax = reinterpret_cast<const uint16_t *>(blob + P.nmax_off);
    const int64_t S = P.stride;      // reads per cell row
    using RC = ReadCursorT<ST>;
    using TAB = TableRef<SMEM, COMPACT>;
    // ST: the tile's cells (W rows of 256 x 8 bytes) live in shared memory behind the blob.  MODE 1: every thread
    // copies its own read's cells there with asynchronous 8-byte copies (no registers, all W in flight at once).
    // MODE 2: they are packed there from the raw bytes in the warps' rings.
    unsigned char *tile_buf = smem_raw + HDR + (SMEM ? P.blob_bytes : 0u);
    const uint32_t cells_bytes = ((uint32_t)P.W * K3_TILE * 8u + 15u) & ~15u;
    uint64_t *s_cells = reinterpret_cast<uint64_t *>(tile_buf + (FUSED ? (uint32_t)group * cells_bytes : 0u));   // [groups][W][256]

    // A CTA takes tiles of 256 reads.  Phase 1 is one thread per read: phasing, filters, and the FIRST anchor
    // step of both junction searches (a wild-type read is finished there).  Searches that are still open
    // become items (read, position) in a shared-memory queue - left items from the front, right items from
    // the back - and are then worked off in rounds by WHICHEVER thread comes next (batch step + anchor step,
    // re-queued while open), so every warp of a round runs one code path with all lanes busy instead of
    // waiting for the slowest read of the warp.  Phase 3 is one thread per read again: junction call, record.
    // (one struct per group: every access is one base register + a constant offset)
    struct GroupState {
        uint32_t left[K3_TILE], right[K3_TILE];     // clj | tlj << 16,  crj | trj << 16
        uint32_t ctx[K3_TILE];                      // stored length | target id << 16
        uint32_t queue[2][2 * K3_TILE];             // item = local read | p << 8
        uint32_t count[3][2];                       // [round % 3][side]
        unsigned int unassigned;                    // fused: aggregation
        uint32_t pad;
    };
    __shared__ GroupState g_state[GMAX];
    GroupState &G = g_state[group];
    uint32_t *s_left = G.left, *s_right = G.right, *s_ctx = G.ctx;
    uint32_t (*s_queue)[2 * K3_TILE] = G.queue;
    uint32_t (*s_count)[2] = G.count;
    unsigned int &s_unassigned = G.unassigned;
    if (tx < 6) s_count[tx >> 1][tx & 1] = 0;

    // ---- fused: accumulators of the aggregation, the warps' raw rings and their mbarriers
    unsigned int *s_hist = nullptr;
    int per_sample = 0;
    ScarTableRef TB; ScarAccRef AC;
    const int warp = tx >> 5, lane = tx & 31;                // warp inside the group
    const int64_t tile_step = (int64_t)gridDim.x * n_groups * K3_TILE;
    uint32_t ring_addr = 0, ring_bar = 0;                    // this warp's slots and their mbarriers
    uint32_t ring_phase[FUSED_SLOTS] = {};                   // staged chunks consumed per slot (mbarrier parity)
    ChunkDesc ring_desc[FUSED_SLOTS];                        // descriptors of the chunks in flight (statically indexed)
    if (FUSED) {
        const uint32_t warp_all = (uint32_t)(group * (K3_TILE / 32) + warp);
        ring_addr = smem_u32(tile_buf) + (uint32_t)n_groups * cells_bytes + warp_all * FUSED_SLOTS * P.raw_cap;
        ring_bar = smem_u32(bar) + 16u + warp_all * FUSED_SLOTS * 8u;
        if (tx == 0) s_unassigned = 0;
        if (lane == 0) {
#pragma unroll
            for (int sl = 0; sl < FUSED_SLOTS; ++sl) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ring_bar + 8u * sl));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        TB.slots = P.slots; TB.slot_mask = P.slot_mask; TB.max_probe = P.max_probe; TB.status = P.status;
        TB.claimed = P.claimed; TB.n_entries = P.n_entries;
        AC.counters = P.counters; AC.phase_hist = P.phase_hist; AC.unassigned = P.unassigned; AC.n_samples = P.n_samples;
        AC.pb = P.phase_bins; AC.joint = P.phase_joint;
        per_sample = acc_words_per_sample(AC.pb, AC.joint);
        s_hist = reinterpret_cast<unsigned int *>(tile_buf + (uint32_t)n_groups * (cells_bytes + (K3_TILE / 32) * FUSED_SLOTS * P.raw_cap));
        if ((P.aggregate & 1) && P.hist_smem)
            for (int i = threadIdx.x; i < P.n_samples * per_sample; i += blockDim.x) s_hist[i] = 0;
        __syncwarp();
        // prologue of the warp's pipeline: the first chunk(s) of its first tile
#pragma unroll
        for (int sl = 0; sl < FUSED_SLOTS; ++sl) {
            ring_desc[sl] = chunk_desc(P.seq, ((int64_t)blockIdx.x * n_groups + group) * K3_TILE + 32 * warp + FUSED_CHUNK * sl, FUSED_CHUNK, P.n_reads, P.raw_cap);
            if (ring_desc[sl].staged && lane == 0) chunk_issue(ring_desc[sl], ring_addr + (uint32_t)sl * P.raw_cap, ring_bar + 8u * sl);
        }
    }
    __syncthreads();
    uint32_t round = 0;
    int epoch = 0;

    for (int64_t tile = ((int64_t)blockIdx.x * n_groups + group) * K3_TILE; tile < P.n_reads; tile += tile_step) {
        const int64_t r = tile + tx;
        bool live = r < P.n_reads;
        const bool in_range = live;
        uint32_t info = 0u;                                     // stored length | N count << 16
        const uint32_t sample = live ? (uint32_t)P.samples[r] : 0xFFFFu;     // (loaded early: used after the pack)
        if (MODE == 1 && live) {
            const uint64_t *src = P.cells + r;
            uint32_t dst = smem_u32(s_cells + tx);
            for (int j = 0; j < P.W; ++j, src += S, dst += K3_TILE * 8)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        if (FUSED) {
            // 16-bit packed counters shared by the CTA's groups: every group empties them (atomic exchange, no barrier)
            // before the reads taken by all groups since anybody's last flush can reach 2^16
            if ((P.aggregate & 1) && P.hist_smem && ++epoch > 65535 / (K3_TILE * n_groups) - 2) {
                acc_flush_exch(AC, s_hist, per_sample, tx, K3_TILE);
                epoch = 1;
            }
            // ---- K1 in shared memory, one pipeline per warp (see pack_chunk): the warp's 32 reads in FUSED_CHUNKS chunks
            __syncwarp();                                 // phase 3 of the previous tile read this warp's cells
            const int64_t w0 = tile + 32 * warp;
            static_assert(!FUSED || FUSED_CHUNKS % FUSED_SLOTS == 0, "a tile's chunks must cycle through the slots evenly");
#pragma unroll 1
            for (int c0 = 0; c0 < FUSED_CHUNKS; c0 += FUSED_SLOTS) {
#pragma unroll
                for (int sl = 0; sl < FUSED_SLOTS; ++sl) {
                    const int c = c0 + sl;
                    const uint32_t slot_addr = ring_addr + (uint32_t)sl * P.raw_cap, slot_bar = ring_bar + 8u * sl;
                    // the descriptor of the refill first: its loads fly while the chunk is packed
                    const int64_t rn = c + FUSED_SLOTS < FUSED_CHUNKS ? w0 + FUSED_CHUNK * (c + FUSED_SLOTS)
                                                                      : w0 + tile_step + FUSED_CHUNK * (c + FUSED_SLOTS - FUSED_CHUNKS);
                    const ChunkDesc dn = chunk_desc(P.seq, rn, FUSED_CHUNK, P.n_reads, P.raw_cap);
                    if (ring_desc[sl].staged) { chunk_wait(slot_bar, ring_phase[sl] & 1u); ++ring_phase[sl]; }
                    pack_chunk<LANES>(P, ring_desc[sl], slot_addr, w0 + FUSED_CHUNK * c, 32 * warp + FUSED_CHUNK * c, s_cells, s_ctx);
                    __syncwarp();                         // every lane has read the slot: refill it FUSED_SLOTS chunks ahead
                    if (dn.staged && lane == 0) chunk_issue(dn, slot_addr, slot_bar);
                    ring_desc[sl] = dn;
#ifdef FUSED_L2_PREFETCH
                    {   // the chunk after the refill goes to L2 now, so that its copy (issued one pack later) is short
                        const int cc = c + 2 * FUSED_SLOTS;
                        const int64_t rp2 = cc < FUSED_CHUNKS ? w0 + FUSED_CHUNK * cc
                                          : (cc < 2 * FUSED_CHUNKS ? w0 + tile_step + FUSED_CHUNK * (cc - FUSED_CHUNKS) : w0 + 2 * tile_step + FUSED_CHUNK * (cc - 2 * FUSED_CHUNKS));
                        const ChunkDesc dp = chunk_desc(P.seq, rp2, FUSED_CHUNK, P.n_reads, P.raw_cap);
                        if (dp.staged && lane == 0) chunk_prefetch_l2(dp);
                    }
#endif
                }
            }
            __syncwarp();
            info = in_range ? s_ctx[tx] : 0u;
        }
        __align__(16) scar_record rec;
        rec.status = SCAR_ST_UNASSIGNED; rec.flags = 0; rec.sample = SCAR_SAMPLE_NONE;
        rec.clj = rec.crj = rec.tlj = rec.trj = 0; rec.hr_n = 0;
        rec.phase_r1 = SCAR_PHASE_NA; rec.phase_r2 = SCAR_PHASE_NA;
        if (sample >= (uint32_t)P.n_samples) live = false;      // unidentified (INDEL_Processing.py:1194-1198)
        else rec.sample = (uint16_t)sample;
        if (!FUSED) info = live ? P.lens[r] : 0u;                // stored length | N count << 16 (from K1)
        if (!live) info = 0u;
        const int len = (int)(info & 0xFFFFu);
        const int tid = live ? __ldg(P.sample_target + sample) : 0;
        const ScarTargetMeta &T = metas[tid];
        const uint64_t *col = P.cells + (in_range ? r : 0);     // cell j of this read = col[j * S]
        const int Wr = (len + 15) >> 4;                          // cells that hold bases
        const int cut = T.cutsite;
        int clj = 0, crj = 0, tlj = cut, trj = cut;
        RC rc0; rc0.col = col; rc0.S = S; rc0.sc = s_cells + tx; rc0.Wr = Wr;
        if (MODE == 1) asm volatile("cp.async.wait_all;" ::: "memory");

        if (live) {
            const int n_count = (int)(info >> 16);
            // ---- phasing (INDEL_Processing.py:948-975): first k matching the prefix / the suffix.
            // The first 16 bases and the last 16 bases are extracted once; a phase is a shift + compare.
            if (P.do_phasing && T.n_phase > 0) {
                const RC &pc0 = rc0;
                uint32_t hcode, hvalid, scode, svalid;
                pc0.word16(0, hcode, hvalid);
                const int sb = len > 16 ? len - 16 : 0;          // suffix window = bases sb .. sb+15
                pc0.word16(sb, scode, svalid);
                int f1 = -1, f2 = -1;
                if (T.ph_lut_off) {
                    // every non-empty phase string of a side has the same length (<= 6): the first matching
                    // k is tabulated per prefix / suffix code at context creation
                    const uint8_t *lut = blob + T.ph_lut_off;
                    const int n1 = T.ph_n1, n2 = T.ph_n2;
                    if (n1 && len >= n1 && (hvalid & ((1u << n1) - 1u)) == ((1u << n1) - 1u))
                        f1 = (int)lut[hcode & ((1u << (2 * n1)) - 1u)] - 1;
                    if (n2 && len >= n2) {
                        const int o = len - n2 - sb;
                        if (((svalid >> o) & ((1u << n2) - 1u)) == ((1u << n2) - 1u))
                            f2 = (int)lut[(1u << (2 * max(n1, n2))) + ((scode >> (2 * o)) & ((1u << (2 * n2)) - 1u))] - 1;
                    }
                } else {
                    for (int k = 0; k < T.n_phase; ++k) {
                        const ScarPhase ph = phases[T.phase_off + k];
                        if (ph.r1_len == 0) continue;            // empty R1 string: phase skipped (:954-957)
                        if (f2 < 0 && ph.r2_len != 0 && ph.r2_len <= 16 && len >= ph.r2_len) {
                            const int o = len - ph.r2_len - sb;      // offset of the suffix inside the window
                            const uint32_t vm = (1u << ph.r2_len) - 1u;
                            const uint32_t cm = ph.r2_len == 16 ? 0xFFFFFFFFu : ((1u << (2 * ph.r2_len)) - 1u);
                            if ((((svalid >> o) & vm) == vm) && (((scode >> (2 * o)) & cm) == ph.r2_code)) f2 = k;
                        }
                        if (f1 < 0 && ph.r1_len <= 16 && len >= ph.r1_len) {
                            const uint32_t vm = (1u << ph.r1_len) - 1u;
                            const uint32_t cm = ph.r1_len == 16 ? 0xFFFFFFFFu : ((1u << (2 * ph.r1_len)) - 1u);
                            if (((hvalid & vm) == vm) && ((hcode & cm) == ph.r1_code)) f1 = k;
                        }
                    }
                }
                rec.phase_r1 = (int8_t)(f1 < 0 ? SCAR_PHASE_NONE : f1);
                rec.phase_r2 = (int8_t)(f2 < 0 ? SCAR_PHASE_NONE : f2);
            }

            // ---- read filters (INDEL_Processing.py:147-154).  nmax[len] is the largest N count with
            // double(n)/double(len) <= N_Limit, tabulated on the host with the same IEEE division.
            if (len == 0 || n_count > (int)nmax[len] || len <= P.min_length) {
                rec.status = SCAR_ST_FILTERED;
                live = false;
            }
        }


        // ---- phase 1: first anchor step of both searches (loci whose 10-mers are all distinct)
        const bool fast = live && T.fast;
        {
            const TAB ptab(blob, P.blob, T.pos);
            LocusRef LR; LR.cut = cut; LR.L = T.length;
            LR.tw = reinterpret_cast<const uint32_t *>(blob + T.tw_off);
            const RC &rc = rc0;
            int p = len - 20;
            bool open = fast && p >= 16;
            if (open) open = left_anchor<SMEM>(rc, ptab, LR, p, clj, tlj);
            __syncwarp();
            queue_push(open, (uint32_t)tx | ((uint32_t)p << 8), &s_queue[round & 1][0], 1, &s_count[round % 3][0]);
            const int plast = len - 26;
            p = 10;
            open = fast && plast >= 10;
            if (open) open = right_anchor<SMEM>(rc, ptab, LR, plast, p, crj, trj);
            __syncwarp();
            queue_push(open, (uint32_t)tx | ((uint32_t)p << 8), &s_queue[round & 1][2 * K3_TILE - 1], -1,
                       &s_count[round % 3][1]);
        }
        bool cutwindow_done = false;          // the exhaustive path checks the cut window itself
        // ---- left junction (SlidingWindow.pyx:49-69): p = len-20 .. 16, descending
        if (live && !T.fast) {
            const TAB ltab(blob, P.blob, T.left);
            const int p0 = len - 20;
            if (p0 >= 16) {
                int j = p0 >> 4;
                uint64_t hi = rc0.cell_or_zero(j + 1);
                uint32_t best = 0;
                for (; j >= 1; --j) {
                    const uint64_t lo = rc0.cell(j);
                    const uint32_t wlo = (uint32_t)lo, whi = (uint32_t)hi;
                    uint32_t m[16];
                    if (probe16(ltab, wlo, whi, m) < 4095u) {
                        // some window of this cell is in the table: keep the first in scan order among
                        // the positions that are scanned (<= p0) and whose window is all ACGT
                        const uint32_t vwin = ((uint32_t)(lo >> 32) & 0xFFFFu) | ((uint32_t)(hi >> 32) << 16);
                        const int pbase = 16 * j, top = p0 - pbase;       // top >= 0 here
                        const uint32_t okmask = runs_of_ten(vwin) & (top >= 15 ? 0xFFFFu : ((2u << top) - 1u));
#pragma unroll
                        for (int o = 15; o >= 0; --o) {
                            const uint32_t cand = ((uint32_t)(pbase + o) << 12) + m[o];
                            if (m[o] < 4095u && (okmask & (1u << o))) best = max(best, cand);
                        }
                        if (best) break;
                    }
                    hi = lo;
                }
                if (best) {
                    const int p = (int)(best >> 12), i = (int)(best & 0xFFFu);
                    clj = p + 10; tlj = cut - i;
                    // SlidingWindow.pyx:56-60 (reachable only through the per-read drop-in)
                    if (T.cutwindow_key != 0xFFFFFFFFu) {
                        const int jj = p >> 4, o = p & 15;
                        const uint64_t a = rc0.cell(jj);
                        const uint64_t b = rc0.cell_or_zero(jj + 1);
                        if ((__funnelshift_r((uint32_t)a, (uint32_t)b, 2 * o) & SCAR_KEY_MASK) == T.cutwindow_key) {
                            rec.status = SCAR_ST_NO_CUT; rec.flags = SCAR_FL_CUTWINDOW;
                            rec.tlj = rec.trj = (uint16_t)cut;
                            live = false;
                        }
                    }
                }
            }
            cutwindow_done = true;
        }
        __syncwarp();

        // ---- right junction (SlidingWindow.pyx:79-93): p = 10 .. len-26, ascending
        if (live && !T.fast) {
            const TAB rtab(blob, P.blob, T.right);
            const int plast = len - 26;
            if (plast >= 10) {
                uint64_t lo = rc0.cell(0);
                uint32_t best = 0xFFFFFFFFu;
                for (int j = 0; 16 * j <= plast; ++j) {
                    const uint64_t hi = rc0.cell_or_zero(j + 1);
                    const uint32_t wlo = (uint32_t)lo, whi = (uint32_t)hi;
                    uint32_t m[16];
                    if (probe16(rtab, wlo, whi, m) < 4095u) {
                        const uint32_t vwin = ((uint32_t)(lo >> 32) & 0xFFFFu) | ((uint32_t)(hi >> 32) << 16);
                        const int pbase = 16 * j, top = plast - pbase;    // top >= 0 here
                        uint32_t okmask = runs_of_ten(vwin) & (top >= 15 ? 0xFFFFu : ((2u << top) - 1u));
                        if (j == 0) okmask &= 0xFC00u;                    // p >= 10
#pragma unroll
                        for (int o = 0; o < 16; ++o) {
                            const uint32_t cand = ((uint32_t)(pbase + o) << 12) + m[o];
                            if (m[o] < 4095u && (okmask & (1u << o))) best = min(best, cand);
                        }
                        if (best != 0xFFFFFFFFu) break;
                    }
                    lo = hi;
                }
                if (best != 0xFFFFFFFFu) { crj = (int)(best >> 12); trj = cut + (int)(best & 0xFFFu); }
            }
        }
        __syncwarp();


        s_left[tx] = (uint32_t)clj | ((uint32_t)tlj << 16);
        s_right[tx] = (uint32_t)crj | ((uint32_t)trj << 16);
        s_ctx[tx] = (uint32_t)len | ((uint32_t)tid << 16);

        // ---- phase 2: rounds over the open searches of the tile
        for (;;) {
            group_sync();
            const uint32_t *q = s_queue[round & 1];
            uint32_t *qn = s_queue[(round + 1) & 1];
            const int nl = (int)s_count[round % 3][0], nr = (int)s_count[round % 3][1];
            if (tx == 0) { s_count[(round + 2) % 3][0] = 0; s_count[(round + 2) % 3][1] = 0; }
            // (the next tile queues into the NEXT slot: a thread that leaves early must not be seen by one that
            // has not read this round's counts yet)
            if (nl + nr == 0) { ++round; break; }
            const int nl32 = (nl + 31) & ~31;                 // warps are pure left or pure right
            for (int v = tx; v < nl32 + ((nr + 31) & ~31); v += K3_TILE) {
                const bool is_left = v < nl32;
                const int k = is_left ? v : v - nl32;
                const bool have = k < (is_left ? nl : nr);
                const uint32_t item = have ? q[is_left ? k : 2 * K3_TILE - 1 - k] : 0u;
                const int rl = (int)(item & 0xFFu);
                int p = (int)(item >> 8);
                const uint32_t ctx = s_ctx[rl];
                const int ilen = (int)(ctx & 0xFFFFu);
                const ScarTargetMeta &IT = metas[ctx >> 16];
                const TAB ptab(blob, P.blob, IT.pos);
                LocusRef LR; LR.cut = IT.cutsite; LR.L = IT.length;
                LR.tw = reinterpret_cast<const uint32_t *>(blob + IT.tw_off);
                RC rc; rc.col = P.cells + tile + rl; rc.S = S; rc.sc = s_cells + rl; rc.Wr = (ilen + 15) >> 4;
                bool open = have;
                if (is_left) {
                    int a = 0, b = LR.cut;
                    if (open) open = left_batch<SMEM>(rc, ptab, LR, p, a, b);
                    __syncwarp();
                    if (open) open = left_anchor<SMEM>(rc, ptab, LR, p, a, b);
                    __syncwarp();
                    if (have && !open && a) s_left[rl] = (uint32_t)a | ((uint32_t)b << 16);
                    queue_push(open, (uint32_t)rl | ((uint32_t)p << 8), qn, 1, &s_count[(round + 1) % 3][0]);
                } else {
                    const int plast = ilen - 26;
                    int a = 0, b = LR.cut;
                    if (open) open = right_batch<SMEM>(rc, ptab, LR, plast, p, a, b);
                    __syncwarp();
                    if (open) open = right_anchor<SMEM>(rc, ptab, LR, plast, p, a, b);
                    __syncwarp();
                    if (have && !open && a) s_right[rl] = (uint32_t)a | ((uint32_t)b << 16);
                    queue_push(open, (uint32_t)rl | ((uint32_t)p << 8), qn + 2 * K3_TILE - 1, -1, &s_count[(round + 1) % 3][1]);
                }
            }
            ++round;
        }

        // ---- phase 3: one thread per read again
        {
            const uint32_t a = s_left[tx], b = s_right[tx];
            clj = (int)(a & 0xFFFFu); tlj = (int)(a >> 16); crj = (int)(b & 0xFFFFu); trj = (int)(b >> 16);
        }
        if (live && !cutwindow_done && clj && T.cutwindow_key != 0xFFFFFFFFu) {   // SlidingWindow.pyx:56-60 (per-read drop-in only)
            uint32_t code, valid;
            rc0.word16(clj - 10, code, valid);
            if ((code & SCAR_KEY_MASK) == T.cutwindow_key) {
                rec.status = SCAR_ST_NO_CUT; rec.flags = SCAR_FL_CUTWINDOW;
                rec.tlj = rec.trj = (uint16_t)cut;
                live = false;
            }
        }

        if (live) {
            rec.clj = (uint16_t)clj; rec.crj = (uint16_t)crj; rec.tlj = (uint16_t)tlj; rec.trj = (uint16_t)trj;
            if (clj < 1 && crj < 1) {                            // :96-98
                rec.status = SCAR_ST_NO_JUNCTION;
                live = false;
            }
        }

        // ---- HR donor scan (:102-117): occurrences at p = 25 .. len-26-h
        if (HR && live) {
            const int h = P.hr_len;
            const int plast = len - 26 - h;
            const uint64_t cm = h == 32 ? ~0ull : ((1ull << (2 * h)) - 1ull);
            const uint64_t vm = h == 32 ? 0xFFFFFFFFull : ((1ull << h) - 1ull);
            int n = 0;
            if (plast >= 25) {
                int j = 25 >> 4;
                uint64_t c0 = rc0.cell(j);
                uint64_t c1 = rc0.cell_or_zero(j + 1);
                uint64_t c2 = rc0.cell_or_zero(j + 2);
                for (int p = 25; p <= plast; ++p) {
                    if ((p >> 4) != j) { j = p >> 4; c0 = c1; c1 = c2; c2 = rc0.cell_or_zero(j + 2); }
                    const int o = p & 15;
                    const uint32_t lo = __funnelshift_r((uint32_t)c0, (uint32_t)c1, 2 * o);
                    const uint32_t hi = __funnelshift_r((uint32_t)c1, (uint32_t)c2, 2 * o);
                    const uint64_t code = (((uint64_t)hi << 32) | lo) & cm;
                    const uint64_t v48 = ((c0 >> 32) & 0xFFFFull) | (((c1 >> 32) & 0xFFFFull) << 16) |
                                         (((c2 >> 32) & 0xFFFFull) << 32);
                    n += ((((v48 >> o) & vm) == vm) && code == P.hr_code) ? 1 : 0;
                }
            }
            rec.hr_n = (uint16_t)n;
            if (n) rec.flags |= SCAR_FL_HR;
        }

        // ---- junction call (:134-171)
        if (live) {
            bool cut_found = false, has_n = false;
            if (0 < clj && clj < crj) {
                // literal 'N' anywhere in read[clj:crj] drops the read with no counter (:140-141)
                for (int j = clj >> 4; 16 * j < crj; ++j) {
                    const int lo = max(clj, 16 * j) - 16 * j, hi = min(crj, 16 * j + 16) - 16 * j;
                    const uint32_t m = ((1u << (hi - lo)) - 1u) << lo;
                    has_n |= (((uint32_t)(rc0.cell(j) >> 48)) & m) != 0;
                }
                if (!has_n) { cut_found = true; rec.flags |= SCAR_FL_INSERTION; }
            }
            if (has_n) rec.status = SCAR_ST_N_INSERTION;
            else {
                if (tlj < cut) { cut_found = true; rec.flags |= SCAR_FL_LEFT_DEL; }
                if (trj > cut) { cut_found = true; rec.flags |= SCAR_FL_RIGHT_DEL; }
                if (clj > crj && crj > 0) { cut_found = true; rec.flags |= SCAR_FL_MICROHOM; }
                rec.status = cut_found ? SCAR_ST_SCAR : SCAR_ST_NO_CUT;
            }
        }
        if (in_range) store_record(P.records + r, rec);

        // ---- fused: stage 4 (INDEL_Processing.py:186-196 and the summary / phasing bookkeeping), what K4 does from
        // the records in a second pass.  The key's segment is hashed from the packed cells in shared memory; only a
        // segment with a non-ACGT byte goes back to the raw read.
        if (FUSED && (P.aggregate & 1)) {
            bool claimed = false;
            uint32_t claimed_slot = 0;
            if (in_range) {
                if (rec.status == SCAR_ST_UNASSIGNED) atomicAdd(&s_unassigned, 1u);
                else if (P.hist_smem) acc_add_smem(s_hist, per_sample, AC.pb, AC.joint, rec);
                else acc_add_global(AC, rec);
                if (rec.status == SCAR_ST_SCAR) {
                    const KeySegment ks = key_segment(rec);
                    uint64_t h1, h2;
                    key_init(ks.hdr, h1, h2);
                    bool acgt = true;
                    for (int q = ks.lo; q < ((P.aggregate & 2) ? ks.lo : ks.hi); q += 32) {
                        const int n = min(32, ks.hi - q);
                        const Word32 w = read32(rc0, q);
                        const uint32_t vm = n == 32 ? 0xFFFFFFFFu : ((1u << n) - 1u);
                        acgt = acgt && (w.valid & vm) == vm;
                        key_step(n == 32 ? w.code : (w.code & ((1ull << (2 * n)) - 1ull)), h1, h2);
                    }
                    if (acgt) key_finish(h1, h2);
                    else {
                        const uint8_t *raw; int rawlen;          // (the fused kernel does not take the Ramsden platform)
                        get_item(P.seq, r, raw, rawlen);
                        uint64_t hh[2];
                        key_hashes_bytes_fwd(ks.hdr, ks.lo, ks.hi, raw + (P.platform == SCAR_PLATFORM_TRUSEQ ? 6 : 0), hh);
                        h1 = hh[0]; h2 = hh[1];
                    }
                    claimed = table_insert(TB, h1, h2, 1u, P.first_index + (uint64_t)r, rec.sample, claimed_slot);
                }
            }
            __syncwarp();
            table_note_claims(TB, claimed, claimed_slot);      // one atomic per warp appends the new slots to the list
        }
    }
    if (FUSED && (P.aggregate & 1)) {
        group_sync();
        if (P.hist_smem) acc_flush_exch(AC, s_hist, per_sample, tx, K3_TILE);      // (a slower group flushes what it adds later)
        if (tx == 0 && s_unassigned) atomicAdd(P.unassigned, (unsigned long long)s_unassigned);
    }
}

// host-side launcher (called from scar_api.cu)
template <int MODE, int LANES>
static void (*pick_kernel(bool smem, bool hr, bool compact))(const K3Params)
{
    if (compact)
        return smem ? (hr ? scar_window_kernel<true, true, MODE, LANES, true> : scar_window_kernel<true, false, MODE, LANES, true>)
                    : (hr ? scar_window_kernel<false, true, MODE, LANES, true> : scar_window_kernel<false, false, MODE, LANES, true>);
    return smem ? (hr ? scar_window_kernel<true, true, MODE, LANES, false> : scar_window_kernel<true, false, MODE, LANES, false>)
                : (hr ? scar_window_kernel<false, true, MODE, LANES, false> : scar_window_kernel<false, false, MODE, LANES, false>);
}

struct K3Plan { void (*kern)(const K3Params); size_t smem; int per_sm; bool stage; int groups = 1; };

static cudaError_t plan_window(const K3Params &P, bool smem_tables, K3Plan &plan)
{
    const int threads = K3_TILE;
    // short reads: the tile's cells are staged in shared memory as well (W rows of 2 KB)
    const bool no_stage = getenv("SCAR_K3_NO_STAGE") != nullptr;
    int stage_max = K3_STAGE_MAX_CELLS;
    if (const char *e = getenv("SCAR_K3_STAGE_MAX_CELLS")) { const int v = atoi(e); if (v >= 1) stage_max = v; }
    const size_t smem_base = 16 + (smem_tables ? P.blob_bytes : 0), cell_bytes = (size_t)P.W * K3_TILE * 8;
    // (227 KB per CTA on sm_100, 7.5 KB of it the kernel's static queues and result arrays)
    plan.stage = P.W <= stage_max && !no_stage && smem_base + cell_bytes + 7680 <= 227 * 1024;
    plan.smem = smem_base + (plan.stage ? cell_bytes : 0);
    const bool hr = P.hr_len > 0;
    plan.kern = plan.stage ? pick_kernel<1, 1>(smem_tables, hr, P.compact_tables != 0) : pick_kernel<0, 1>(smem_tables, hr, P.compact_tables != 0);
    cudaError_t e = cudaFuncSetAttribute(plan.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    if (e != cudaSuccess) return e;
    plan.per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&plan.per_sm, plan.kern, threads, plan.smem);
    if (e != cudaSuccess) return e;
    if (plan.per_sm < 1) plan.per_sm = 1;
    if (const char *e2 = getenv("SCAR_K3_CTAS_PER_SM")) { const int v = atoi(e2); if (v >= 1 && v < plan.per_sm) plan.per_sm = v; }
    return cudaSuccess;
}

// how K3 would be launched for this problem: cells staged in shared memory?, resident CTAs per SM, dynamic smem bytes
cudaError_t scar_window_plan(const K3Params &P, bool smem_tables, int *stage, int *per_sm, int *smem_bytes)
{
    K3Plan plan;
    cudaError_t e = plan_window(P, smem_tables, plan);
    if (e != cudaSuccess) return e;
    if (stage) *stage = plan.stage ? 1 : 0;
    if (per_sm) *per_sm = plan.per_sm;
    if (smem_bytes) *smem_bytes = (int)plan.smem;
    return cudaSuccess;
}

cudaError_t scar_launch_window(const K3Params &P, bool smem_tables, int sm_count, cudaStream_t stream,
                               int *launches)
{
    if (P.n_reads <= 0) return cudaSuccess;
    const int threads = K3_TILE;
    K3Plan plan;
    cudaError_t e = plan_window(P, smem_tables, plan);
    if (e != cudaSuccess) return e;
    // persistent grid: a whole number of resident CTAs per SM, never more than the work needs
    int64_t want = (P.n_reads + threads - 1) / threads;
    int64_t grid = (int64_t)sm_count * plan.per_sm;
    if (grid > want) grid = want;
    plan.kern<<<(unsigned)grid, threads, plan.smem, stream>>>(P);
    if (launches) ++*launches;
    return cudaGetLastError();
}

// ---- fused K1 + K3 (+ K4) from raw bytes ----------------------------------------------------------------
bool scar_fused_eligible(int W, int platform)
{
    // reads up to 320 nt (20 cells held in registers while the raw bytes are overwritten); the Ramsden platform
    // stores the reverse complement of the read and stays on the K1 -> K3 -> K4 path
    return W <= 20 && platform != SCAR_PLATFORM_RAMSDEN && getenv("SCAR_NO_FUSE") == nullptr;
}

static cudaError_t plan_fused(K3Params &P, bool smem_tables, int max_read_len, K3Plan &plan)
{
    const int threads = K3_TILE, warps = K3_TILE / 32;
    // one ring slot holds a chunk of 32 / lanes fixed-stride reads of the longest length (+ alignment, + 32 bytes of
    // slack for word loads past a read's end); wider spans (FASTQ views) are read from global memory unless the
    // caller asks for larger slots (SCAR_FUSED_SLOT_BYTES)
    const int raw_len = max_read_len + (P.platform == SCAR_PLATFORM_TRUSEQ ? 12 : 0);
    // lanes per read: 2 (16-read chunks) for reads up to 160 nt, 4 (8-read chunks: half the ring) for longer reads
#ifndef FUSED_LANES_SHORT
#define FUSED_LANES_SHORT 2
#endif
    const int lanes = P.W <= 10 ? FUSED_LANES_SHORT : 4, chunk = 32 / lanes, gmax = fused_groups_max(lanes);
    const size_t cells_bytes = ((size_t)P.W * threads * 8 + 15) & ~(size_t)15;
    const size_t fixed = 16 + (size_t)gmax * warps * FUSED_SLOTS * 8 + (smem_tables ? P.blob_bytes : 0);
    const bool hr = P.hr_len > 0;
    const bool compact = P.compact_tables != 0;
#if FUSED_LANES_SHORT == 1
    if (lanes == 1) plan.kern = pick_kernel<2, 1>(smem_tables, hr, compact);
    else
#endif
    plan.kern = lanes == 2 ? pick_kernel<2, 2>(smem_tables, hr, compact) : pick_kernel<2, 4>(smem_tables, hr, compact);
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, plan.kern);
    if (e != cudaSuccess) return e;
    // what one CTA may use: 227 KB minus the kernel's static arrays (queues and result arrays of gmax groups)
    const size_t budget = 227 * 1024 - fa.sharedSizeBytes - 1024;
    // Ring slot: a chunk of fixed-stride reads of the longest length (+ alignment, + 32 bytes of slack for word loads
    // past a read's end); FASTQ views get slots wide enough for whole records when that still leaves room for two
    // groups (a chunk whose span does not fit its slot is read from global memory).
    const uint32_t slot_reads = ((uint32_t)chunk * (uint32_t)raw_len + 16u + 32u + 15u) & ~15u;
    uint32_t slot = slot_reads;
    if (P.seq.offsets && P.seq.lengths) {
        const uint32_t slot_records = ((uint32_t)chunk * (2u * (uint32_t)raw_len + 64u) + 48u + 15u) & ~15u;
        if (fixed + 2 * (cells_bytes + (size_t)warps * FUSED_SLOTS * slot_records) + 16 * 1024 <= budget) slot = slot_records;
    }
    if (const char *e3 = getenv("SCAR_FUSED_SLOT_BYTES")) { const long v = atol(e3); if (v >= 64) slot = ((uint32_t)v + 15u) & ~15u; }
    P.raw_cap = slot;
    const size_t per_group = cells_bytes + (size_t)warps * FUSED_SLOTS * slot;
    // accumulators: joint phase histogram when it is small, else two marginals; global atomics when neither fits
    P.hist_smem = 0;
    size_t hist = 0;
    if ((P.aggregate & 1)) {
        for (int joint = P.phase_bins * P.phase_bins <= 32 ? 1 : 0; joint >= 0 && !P.hist_smem; --joint) {
            const size_t h = (size_t)P.n_samples * acc_words_per_sample(P.phase_bins, joint) * sizeof(unsigned int);
            if (h <= 48 * 1024 && fixed + per_group + h <= budget) { P.hist_smem = 1; P.phase_joint = joint; hist = h; }
        }
    }
    if (fixed + per_group + hist > budget) return cudaErrorInvalidConfiguration;     // the caller takes the three-kernel path
    // groups of 256 threads per CTA: as many as fit beside the shared blob and accumulators
    int groups = gmax;
    while (groups > 1 && fixed + (size_t)groups * per_group + hist > budget) --groups;
    if (const char *e2 = getenv("SCAR_FUSED_GROUPS")) { const int v = atoi(e2); if (v >= 1 && v < groups) groups = v; }
    plan.stage = true;
    plan.groups = groups;
    plan.smem = fixed + (size_t)groups * per_group + hist;
    e = cudaFuncSetAttribute(plan.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    if (e != cudaSuccess) return e;
    plan.per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&plan.per_sm, plan.kern, threads * groups, plan.smem);
    if (e != cudaSuccess) return e;
    if (plan.per_sm < 1) plan.per_sm = 1;
    return cudaSuccess;
}

cudaError_t scar_fused_plan(K3Params P, bool smem_tables, int max_read_len, int *per_sm, int *smem_bytes)
{
    K3Plan plan;
    cudaError_t e = plan_fused(P, smem_tables, max_read_len, plan);
    if (e != cudaSuccess) return e;
    if (per_sm) *per_sm = plan.per_sm * plan.groups;       // 256-thread groups resident per SM
    if (smem_bytes) *smem_bytes = (int)plan.smem;
    return cudaSuccess;
}

cudaError_t scar_launch_fused(K3Params P, bool smem_tables, int max_read_len, int sm_count, cudaStream_t stream, int *launches)
{
    if (P.n_reads <= 0) return cudaSuccess;
    const int threads = K3_TILE;
    K3Plan plan;
    cudaError_t e = plan_fused(P, smem_tables, max_read_len, plan);
    if (e != cudaSuccess) return e;
    const int64_t tiles = (P.n_reads + threads - 1) / threads;
    int64_t want = (tiles + plan.groups - 1) / plan.groups;
    int64_t grid = (int64_t)sm_count * plan.per_sm;
    if (grid > want) grid = want;
    plan.kern<<<(unsigned)grid, threads * plan.groups, plan.smem, stream>>>(P);
    if (launches) ++*launches;
    return cudaGetLastError();
}
