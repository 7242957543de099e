// scar_homeology_core.h — the two per-unique-scar searches of ScarSearch.frequency_output, one key at a time:
//   homeology_search            reference scarmapper/INDEL_Processing.py:550-657
//   templated_insertion_search  reference scarmapper/INDEL_Processing.py:659-700
// Host + device (the kernel in scar_homeology.cu maps one thread to one key; tests/test_homeology_host.py
// compiles this header for the host and compares it with the reference's own methods).
//
// Both searches are Python string slicing + 5-mer comparisons; what has to be reproduced exactly is the
// slice arithmetic (negative starts wrap around, everything is clipped to the string) and
// match_maker(query, unknown) = Levenshtein(query, unknown) - (len(unknown) - len(query))
// (Valkyries/Sequence_Magic.py:123-136) on strings of at most 5 characters.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SCAR_HOM __host__ __device__ __forceinline__
#else
#define SCAR_HOM static inline
#endif

// results of one key, all positions relative to the strings passed in (see scar_homeology_batch)
struct ScarHomeology {
    int32_t l_dist, l_pos, l_q_start, l_q_len, l_s_start, l_s_len;   // l_dist: 0 / 1, or -1 = nothing found
    int32_t r_dist, r_pos, r_q_start, r_q_len, r_s_start, r_s_len;
    int32_t lt_start, rt_start;                                      // 5-nt templates in the target region, -1 = none
    int32_t pad0, pad1;
};

struct HomSlice { int32_t start, len; };

// s[a:b] of a string of length n, as Python evaluates it
SCAR_HOM HomSlice hom_pyslice(int32_t n, int32_t a, int32_t b)
{
    if (a < 0) { a += n; if (a < 0) a = 0; } else if (a > n) a = n;
    if (b < 0) { b += n; if (b < 0) b = 0; } else if (b > n) b = n;
    HomSlice s; s.start = a; s.len = b > a ? b - a : 0;
    return s;
}

// unit-cost edit distance of two strings of at most 5 characters
SCAR_HOM int hom_lev5(const uint8_t *a, int na, const uint8_t *b, int nb)
{
    int row[6];
    for (int j = 0; j <= nb; ++j) row[j] = j;
    for (int i = 1; i <= na; ++i) {
        int diag = row[0];
        row[0] = i;
        for (int j = 1; j <= nb; ++j) {
            const int up = row[j];
            int v = diag + (a[i - 1] != b[j - 1] ? 1 : 0);
            if (up + 1 < v) v = up + 1;
            if (row[j - 1] + 1 < v) v = row[j - 1] + 1;
            diag = up;
            row[j] = v;
        }
    }
    return row[nb];
}

// Sequence_Magic.rcomp's translate table (Valkyries/Sequence_Magic.py:97-118); unmapped bytes pass
SCAR_HOM uint8_t hom_comp(uint8_t c)
{
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

// homeology_search(consensus, clj, crj, target, insertion, tlj, trj); only len(insertion) matters
SCAR_HOM void hom_homeology(const uint8_t *cons, int32_t n_cons, int32_t clj, int32_t crj, const uint8_t *target,
                            int32_t n_target, int32_t n_ins, int32_t tlj, int32_t trj, ScarHomeology &out)
{
    const int target_size = 5;
    int junction_padding = target_size;
    if (trj - tlj < target_size) junction_padding = trj - tlj;
    const HomSlice lq = hom_pyslice(n_cons, clj - target_size + n_ins, clj + n_ins);
    const HomSlice rq = hom_pyslice(n_cons, crj - n_ins, crj - n_ins + target_size);
    out.l_dist = -1; out.l_pos = 0; out.l_q_start = lq.start; out.l_q_len = lq.len; out.l_s_start = 0; out.l_s_len = 0;
    out.r_dist = -1; out.r_pos = 0; out.r_q_start = rq.start; out.r_q_len = rq.len; out.r_s_start = 0; out.r_s_len = 0;

    // from the left junction (:601-621)
    int pos = trj - junction_padding;
    bool homology = false, homeology = false;
    for (int it = 0; it < target_size && !homology; ++it) {
        const HomSlice seg = hom_pyslice(n_target, pos, pos + target_size);
        // match_maker(target_segment, lft_query): query = the segment, unknown = lft_query
        const int d = hom_lev5(target + seg.start, seg.len, cons + lq.start, lq.len) - (lq.len - seg.len);
        if (d == 0) {
            out.l_dist = 0; out.l_pos = it - junction_padding; out.l_s_start = seg.start; out.l_s_len = seg.len;
            homology = true;
        } else if (d == 1 && !homeology && target[seg.start + seg.len - 1] == cons[lq.start + lq.len - 1]) {
            // (d == 1 implies that neither string is empty)
            out.l_dist = 1; out.l_pos = it - junction_padding; out.l_s_start = seg.start; out.l_s_len = seg.len;
            homeology = true;
        }
        ++pos;
        if (homology) ++pos;
    }
    // from the right junction (:623-645)
    pos = tlj + junction_padding;
    homology = false; homeology = false;
    for (int it = 0; it < target_size && !homology; ++it) {
        const HomSlice seg = hom_pyslice(n_target, pos - target_size, pos);
        const int d = hom_lev5(target + seg.start, seg.len, cons + rq.start, rq.len) - (rq.len - seg.len);
        if (d == 0) {
            out.r_dist = 0; out.r_pos = it - junction_padding; out.r_s_start = seg.start; out.r_s_len = seg.len;
            homology = true;
        } else if (d == 1 && !homeology && target[seg.start] == cons[rq.start]) {
            out.r_dist = 1; out.r_pos = it - junction_padding; out.r_s_start = seg.start; out.r_s_len = seg.len;
            homeology = true;
        }
        --pos;
        if (homology) --pos;
    }
}

// does region[s.start : +s.len] equal the 5 bytes q?
SCAR_HOM bool hom_eq5(const uint8_t *region, HomSlice s, const uint8_t *q, int nq)
{
    if (s.len != nq) return false;
    for (int k = 0; k < nq; ++k) if (region[s.start + k] != q[k]) return false;
    return true;
}

// templated_insertion_search(insertion, tlj, trj): the first 5-nt stretch of the target region, walking away
// from each junction for at most 50 positions, that equals one end of the insertion or its reverse complement
SCAR_HOM void hom_templates(const uint8_t *ins, int32_t n_ins, const uint8_t *region, int32_t n_region, int32_t tlj,
                            int32_t trj, ScarHomeology &out)
{
    out.lt_start = -1; out.rt_start = -1;
    // insertion[:5], insertion[-5:] and their reverse complements
    uint8_t head[5], tail[5], head_rc[5], tail_rc[5];
    const HomSlice hs = hom_pyslice(n_ins, 0, 5), ts = hom_pyslice(n_ins, -5, n_ins);
    for (int k = 0; k < hs.len; ++k) head[k] = ins[hs.start + k];
    for (int k = 0; k < ts.len; ++k) tail[k] = ins[ts.start + k];
    for (int k = 0; k < hs.len; ++k) head_rc[k] = hom_comp(head[hs.len - 1 - k]);
    for (int k = 0; k < ts.len; ++k) tail_rc[k] = hom_comp(tail[ts.len - 1 - k]);
    const int lower = tlj - 50, upper = trj + 50;
    int lp = tlj - 5, rp = tlj;
    while (rp > lower) {                       // lft_query1 = rcomp(insertion[:5]), lft_query2 = insertion[-5:]
        const HomSlice seg = hom_pyslice(n_region, lp, rp);
        if (hom_eq5(region, seg, head_rc, hs.len) || hom_eq5(region, seg, tail, ts.len)) { out.lt_start = seg.start; break; }
        --lp; --rp;
    }
    lp = trj; rp = trj + 5;
    while (lp < upper) {                       // rt_query1 = insertion[:5], rt_query2 = rcomp(insertion[-5:])
        const HomSlice seg = hom_pyslice(n_region, lp, rp);
        if (hom_eq5(region, seg, head, hs.len) || hom_eq5(region, seg, tail_rc, ts.len)) { out.rt_start = seg.start; break; }
        ++lp; ++rp;
    }
}
