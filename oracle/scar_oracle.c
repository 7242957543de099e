/*
 * scar_oracle.c — CPU restatement of ScarMapper's per-read scar-classification path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under scarmapper_b200/ may import, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the reported CPU baseline.
 *
 * Parity status:
 *   - sliding_window, window_mapping, cutsite_search, filters, phasing: PINNED against the
 *     reference itself (compiled scarmapper/SlidingWindow.pyx and the imported Python driver),
 *     see oracle/build_ref.py, tests/golden/make_golden.py and tests/test_oracle_vs_reference.py.
 *   - Levenshtein distance: python-Levenshtein (requirements.txt:12, "=0.12.2") is NOT under
 *     /root/reference and not installed; the reference holds no tests or golden vectors for it.
 *     "parity unpinned" at that boundary: orc_levenshtein restates the published unit-cost
 *     (insert = delete = substitute = 1) distance that Levenshtein.distance documents.
 *
 * Every function is written the slow, literal way (the loops the reference runs), not the way
 * the CUDA kernels compute the same thing.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/scar_records.h"

/* ------------------------------------------------------------------------------------------
 * Levenshtein.distance(a, b) — unit-cost edit distance, full (m+1)x(n+1) DP, two rows.
 * Call sites in the reference: Valkyries/Sequence_Magic.py:9,131.
 * ---------------------------------------------------------------------------------------- */
int orc_levenshtein(const uint8_t *a, int la, const uint8_t *b, int lb)
{
    int *prev = (int *)malloc(sizeof(int) * (size_t)(lb + 1));
    int *cur = (int *)malloc(sizeof(int) * (size_t)(lb + 1));
    for (int j = 0; j <= lb; ++j) prev[j] = j;
    for (int i = 1; i <= la; ++i) {
        cur[0] = i;
        for (int j = 1; j <= lb; ++j) {
            int sub = prev[j - 1] + (a[i - 1] != b[j - 1]);
            int del = prev[j] + 1;
            int ins = cur[j - 1] + 1;
            int m = sub < del ? sub : del;
            cur[j] = m < ins ? m : ins;
        }
        int *t = prev; prev = cur; cur = t;
    }
    int d = prev[lb];
    free(prev); free(cur);
    return d;
}

/* Sequence_Magic.match_maker (Valkyries/Sequence_Magic.py:123-136):
 *   distance(query, unknown) - (len(unknown) - len(query)) */
int orc_match_maker(const uint8_t *query, int lq, const uint8_t *unknown, int lu)
{
    return orc_levenshtein(query, lq, unknown, lu) - (lu - lq);
}

/* Sequence_Magic.rcomp (Valkyries/Sequence_Magic.py:52-120): reverse, then str.translate with
 * the ambiguous-DNA complement table (upper and lower case); unmapped characters pass through. */
static uint8_t comp_byte(uint8_t c)
{
    switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'M': return 'K'; case 'R': return 'Y'; case 'W': return 'W'; case 'S': return 'S';
    case 'Y': return 'R'; case 'K': return 'M'; case 'V': return 'B'; case 'H': return 'D';
    case 'D': return 'H'; case 'B': return 'V'; case 'X': return 'X'; case 'N': return 'N';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    case 'm': return 'k'; case 'r': return 'y'; case 'w': return 'w'; case 's': return 's';
    case 'y': return 'r'; case 'k': return 'm'; case 'v': return 'b'; case 'h': return 'd';
    case 'd': return 'h'; case 'b': return 'v'; case 'x': return 'x'; case 'n': return 'n';
    default: return c;
    }
}

void orc_rcomp(const uint8_t *s, int n, uint8_t *out)
{
    for (int i = 0; i < n; ++i) out[i] = comp_byte(s[n - 1 - i]);
}

static int slice_eq(const uint8_t *a, int la, const uint8_t *b, int lb)
{
    return la == lb && (la == 0 || memcmp(a, b, (size_t)la) == 0);
}

/* ScarSearch.cutsite_search (scarmapper/INDEL_Processing.py:759-795).
 * Returns the cutsite or -1 when the sgRNA does not map (the reference raises SystemExit). */
int orc_cutsite_search(const uint8_t *target, int L, const uint8_t *sgrna, int ls, int reverse_comp)
{
    uint8_t *work = (uint8_t *)malloc((size_t)(ls > 0 ? ls : 1));
    if (reverse_comp) orc_rcomp(sgrna, ls, work); else memcpy(work, sgrna, (size_t)ls);
    int lft = 0, rt = ls, upper = L - 1, cutsite = -1;
    while (cutsite < 0 && rt < upper) {
        /* target_region[lft:rt] is always ls long here because rt < L-1 */
        if (slice_eq(target + lft, rt - lft, work, ls))
            cutsite = reverse_comp ? lft + 3 : rt - 3;
        ++lft; ++rt;
    }
    free(work);
    return cutsite;
}

/* ScarSearch.window_mapping (scarmapper/INDEL_Processing.py:57-80): number of left/right
 * 10-nt target windows.  Left window i  = target[cut-10-i : cut-i],  i = 0 .. n_left-1
 *                        Right window i = target[cut+i : cut+i+10],  i = 0 .. n_right-1 */
void orc_window_counts(int L, int cutsite, int *n_left, int *n_right)
{
    int nl = 0, nr = 0;
    int rt = cutsite;
    while (rt > 15) { ++nl; --rt; }                 /* lower_limit = 15 */
    int lft = cutsite;
    while (lft < L - 15) { ++nr; ++lft; }           /* upper_limit = L - 15 */
    *n_left = nl; *n_right = nr;
}

/* Python slice target[a:b] for 0 <= a: clamp to the string. */
static int py_slice(int L, int a, int b, int *start)
{
    if (a < 0) a = 0;            /* callers never pass negatives; keep it total */
    if (b > L) b = L;
    if (a > L) a = L;
    *start = a;
    return b > a ? b - a : 0;
}

/* SlidingWindow.sliding_window (scarmapper/SlidingWindow.pyx:20-171), literal loops.
 * The window lists are re-derived from (target, cutsite) exactly as window_mapping builds them.
 * rec->sample / phase fields are left untouched. */
void orc_sliding_window(const uint8_t *read, int len, const uint8_t *target, int L, int cutsite,
                        int lower_limit, const uint8_t *cutwindow, int cwlen,
                        const uint8_t *hr_donor, int hlen, scar_record *rec)
{
    int n_left, n_right;
    orc_window_counts(L, cutsite, &n_left, &n_right);

    int clj = 0, crj = 0, tlj = cutsite, trj = cutsite;
    int left_found = 0, right_found = 0;
    rec->flags = 0; rec->hr_n = 0;

    /* :49-69 left junction */
    int rt_pos = len - 10, lft_pos = rt_pos - 10;
    while (!left_found && lft_pos > lower_limit) {
        const uint8_t *q = read + lft_pos;     /* always 10 long: 15 < lft_pos, rt_pos <= len-10 */
        for (int i = 0; i < n_left; ++i) {
            int s, wl = py_slice(L, cutsite - 10 - i, cutsite - i, &s);
            if (slice_eq(q, 10, target + s, wl)) {
                if (slice_eq(q, 10, cutwindow, cwlen)) {           /* :56-60 */
                    rec->status = SCAR_ST_NO_CUT; rec->flags = SCAR_FL_CUTWINDOW;
                    rec->clj = 0; rec->crj = 0; rec->tlj = (uint16_t)cutsite; rec->trj = (uint16_t)cutsite;
                    return;
                }
                left_found = 1; tlj = cutsite - i; clj = rt_pos;
                break;
            }
        }
        --lft_pos; --rt_pos;
    }

    /* :79-93 right junction; upper_consensus_limit = len - 15 */
    lft_pos = 10; rt_pos = 20;
    while (!right_found && rt_pos < len - 15) {
        const uint8_t *q = read + lft_pos;
        for (int i = 0; i < n_right; ++i) {
            int s, wl = py_slice(L, cutsite + i, cutsite + i + 10, &s);
            if (slice_eq(q, 10, target + s, wl)) {
                right_found = 1; trj = cutsite + i; crj = lft_pos;
                break;
            }
        }
        ++lft_pos; ++rt_pos;
    }

    rec->clj = (uint16_t)clj; rec->crj = (uint16_t)crj; rec->tlj = (uint16_t)tlj; rec->trj = (uint16_t)trj;

    /* :96-98 */
    if (clj < 1 && crj < 1) { rec->status = SCAR_ST_NO_JUNCTION; return; }

    /* :102-117 HR donor scan */
    if (hlen > 0) {
        int r = hlen + 25, l = 25, n = 0;
        while (r < len - 25) {
            if (slice_eq(read + l, r - l, hr_donor, hlen)) ++n;
            ++r; ++l;
        }
        rec->hr_n = (uint16_t)n;
        if (n) rec->flags |= SCAR_FL_HR;
    }

    int cut_found = 0;
    /* :136-145 insertion */
    if (0 < clj && clj < crj) {
        for (int p = clj; p < crj; ++p)
            if (read[p] == 'N') { rec->status = SCAR_ST_N_INSERTION; return; }
        cut_found = 1; rec->flags |= SCAR_FL_INSERTION;
    }
    if (tlj < cutsite) { cut_found = 1; rec->flags |= SCAR_FL_LEFT_DEL; }   /* :148-150 */
    if (trj > cutsite) { cut_found = 1; rec->flags |= SCAR_FL_RIGHT_DEL; }  /* :153-155 */
    if (clj > crj && crj > 0) { cut_found = 1; rec->flags |= SCAR_FL_MICROHOM; } /* :159-163 */

    rec->status = cut_found ? SCAR_ST_SCAR : SCAR_ST_NO_CUT;                /* :166-171 */
}

/* Python seq[-n:] -> (start, length);  seq[:n] is (0, min(n, len)). */
static int py_suffix(int len, int n, int *start)
{
    if (n == 0) { *start = 0; return len; }       /* seq[-0:] is the whole string */
    if (n >= len) { *start = 0; return len; }
    *start = len - n; return n;
}

/* DataProcessing.index_matching (scarmapper/INDEL_Processing.py:1126-1201), PEAR mode
 * (fastq2_read is None, so the loop breaks on the first accepted index, :1185-1186).
 *   left_index[s]  = index_dict[s][0] = Master_Index_File column 3, upper-cased (:1110-1113)
 *   right_index[s] = index_dict[s][2] = Master_Index_File column 2, upper-cased
 *   platform 0 Illumina (mismatch 1): fields f0,f1 = name.split(":")[-1].split("+")[0..1]
 *   platform 1 TruSeq  (mismatch 0), platform 2 Ramsden (mismatch 3).
 * Returns the manifest position of the accepted sample or -1. */
int orc_index_matching(int platform, int mismatch, int n_samples,
                       const uint8_t *left_bytes, const int64_t *left_off,
                       const uint8_t *right_bytes, const int64_t *right_off,
                       const uint8_t *seq, int len,
                       const uint8_t *f0, int l0, const uint8_t *f1, int l1)
{
    for (int s = 0; s < n_samples; ++s) {
        const uint8_t *li = left_bytes + left_off[s];  int ll = (int)(left_off[s + 1] - left_off[s]);
        const uint8_t *ri = right_bytes + right_off[s]; int rl = (int)(right_off[s + 1] - right_off[s]);
        int left_match = 5, right_match = 5;
        if (platform == 0) {                                   /* :1153-1156 */
            right_match = orc_match_maker(ri, rl, f0, l0);
            left_match = orc_match_maker(li, ll, f1, l1);
        } else if (platform == 1) {                            /* :1158-1161 */
            int st, n = py_suffix(len, 6, &st);
            left_match = orc_match_maker(ri, rl, seq, len < 6 ? len : 6);
            right_match = orc_match_maker(li, ll, seq + st, n);
        } else {                                               /* :1163-1171, PEAR */
            int st, n = py_suffix(len, ll, &st);
            left_match = orc_match_maker(li, ll, seq + st, n);
            right_match = orc_match_maker(ri, rl, seq, len < rl ? len : rl);
        }
        if (left_match <= mismatch && right_match <= mismatch) return s;   /* :1176 */
    }
    return -1;
}

/* Phasing scan of consensus_demultiplex (scarmapper/INDEL_Processing.py:948-975).
 * r1[k], r2[k] are the TargetMapper.phasing strings of the read's locus (TargetMapper.py:54-69).
 * out[0] = first k with r1[k] == seq[:len(r1[k])], out[1] = first k with r2[k] == rcomp(seq[-len(r2[k]):]);
 * SCAR_PHASE_NONE when no k matches; phases whose R1 string is empty are skipped (:954-957). */
void orc_phasing(const uint8_t *seq, int len, int n_phase,
                 const uint8_t *r1_bytes, const int64_t *r1_off,
                 const uint8_t *r2_bytes, const int64_t *r2_off, int8_t *out)
{
    int r1_found = -1, r2_found = -1;
    uint8_t *tmp = (uint8_t *)malloc((size_t)(len > 0 ? len : 1));
    for (int k = 0; k < n_phase; ++k) {
        const uint8_t *p1 = r1_bytes + r1_off[k]; int n1 = (int)(r1_off[k + 1] - r1_off[k]);
        const uint8_t *p2 = r2_bytes + r2_off[k]; int n2 = (int)(r2_off[k + 1] - r2_off[k]);
        if (n1 == 0) continue;
        if (r2_found < 0) {
            int st, n = py_suffix(len, n2, &st);
            orc_rcomp(seq + st, n, tmp);
            if (slice_eq(p2, n2, tmp, n)) r2_found = k;
        }
        if (r1_found < 0) {
            int n = len < n1 ? len : n1;
            if (slice_eq(p1, n1, seq, n)) r1_found = k;
        }
    }
    free(tmp);
    out[0] = (int8_t)(r1_found < 0 ? SCAR_PHASE_NONE : r1_found);
    out[1] = (int8_t)(r2_found < 0 ? SCAR_PHASE_NONE : r2_found);
}

/* Read filters of ScarSearch.data_processing (scarmapper/INDEL_Processing.py:147-154):
 *   seq.count("N") / len(seq) > N_Limit  (Python true division = IEEE double divide)
 *   len(seq) <= Minimum_Length
 * len == 0 raises ZeroDivisionError in the reference; reported as filtered here. */
int orc_is_filtered(const uint8_t *seq, int len, double n_limit, int min_length)
{
    if (len == 0) return 1;
    int n = 0;
    for (int i = 0; i < len; ++i) n += seq[i] == 'N';
    if ((double)n / (double)len > n_limit) return 1;
    if (len <= min_length) return 1;
    return 0;
}

/* Platform transform of the stored sequence (INDEL_Processing.py:946,978,981,1181-1182):
 * Illumina: as is; TruSeq: seq[6:-6]; Ramsden: rcomp(seq).  Returns the new length. */
int orc_platform_seq(int platform, const uint8_t *seq, int len, uint8_t *out)
{
    if (platform == 1) {
        int a = len < 6 ? len : 6;
        int b = len - 6; if (b < 0) b = 0;   /* seq[6:-6]: stop index len-6 clamped at 0 */
        int n = b > a ? b - a : 0;
        memcpy(out, seq + a, (size_t)n);
        return n;
    }
    if (platform == 2) { orc_rcomp(seq, len, out); return len; }
    memcpy(out, seq, (size_t)len);
    return len;
}

/* Whole hot path over a batch, one read after the other, the way main_loop drives it
 * (INDEL_Processing.py:1038-1061): demultiplex -> phasing -> platform transform -> filters ->
 * sliding_window.  Aggregation into the scar table is done by the Python wrapper with a dict,
 * like the reference (:186-196).
 *
 * seq/idx CSR inputs; per-sample target id; per-target sequences, cutsites and phasing lists.
 * sample_in: if non-NULL the demultiplex stage is skipped and sample_in[r] is used (-1 = none). */
void orc_classify_batch(
    int64_t n_reads, const uint8_t *seq_bytes, const int64_t *seq_off,
    const uint8_t *f0_bytes, const int64_t *f0_off, const uint8_t *f1_bytes, const int64_t *f1_off,
    const int32_t *sample_in,
    int platform, int mismatch, int n_samples,
    const uint8_t *left_bytes, const int64_t *left_off,
    const uint8_t *right_bytes, const int64_t *right_off,
    const int32_t *sample_target,
    int n_targets, const uint8_t *target_bytes, const int64_t *target_off, const int32_t *cutsite,
    const int32_t *phase_off, /* [n_targets+1] into the r1/r2 CSR lists */
    const uint8_t *r1_bytes, const int64_t *r1_off, const uint8_t *r2_bytes, const int64_t *r2_off,
    const uint8_t *hr_donor, int hlen, double n_limit, int min_length,
    scar_record *out)
{
    (void)n_targets;
    int maxlen = 1;
    for (int64_t r = 0; r < n_reads; ++r) {
        int l = (int)(seq_off[r + 1] - seq_off[r]);
        if (l > maxlen) maxlen = l;
    }
    uint8_t *buf = (uint8_t *)malloc((size_t)maxlen);
    for (int64_t r = 0; r < n_reads; ++r) {
        const uint8_t *seq = seq_bytes + seq_off[r];
        int len = (int)(seq_off[r + 1] - seq_off[r]);
        scar_record *rec = out + r;
        memset(rec, 0, sizeof(*rec));
        rec->phase_r1 = SCAR_PHASE_NA; rec->phase_r2 = SCAR_PHASE_NA;
        int s;
        if (sample_in) s = sample_in[r];
        else {
            const uint8_t *f0 = 0, *f1 = 0; int l0 = 0, l1 = 0;
            if (f0_bytes) { f0 = f0_bytes + f0_off[r]; l0 = (int)(f0_off[r + 1] - f0_off[r]); }
            if (f1_bytes) { f1 = f1_bytes + f1_off[r]; l1 = (int)(f1_off[r + 1] - f1_off[r]); }
            s = orc_index_matching(platform, mismatch, n_samples, left_bytes, left_off,
                                   right_bytes, right_off, seq, len, f0, l0, f1, l1);
        }
        if (s < 0) { rec->status = SCAR_ST_UNASSIGNED; rec->sample = SCAR_SAMPLE_NONE; continue; }
        rec->sample = (uint16_t)s;
        int t = sample_target[s];
        if (platform == 0 && phase_off) {
            int k0 = phase_off[t], nk = phase_off[t + 1] - k0;
            if (nk > 0) {
                int8_t ph[2];
                orc_phasing(seq, len, nk, r1_bytes, r1_off + k0, r2_bytes, r2_off + k0, ph);
                rec->phase_r1 = ph[0]; rec->phase_r2 = ph[1];
            }
        }
        int n = orc_platform_seq(platform, seq, len, buf);
        if (orc_is_filtered(buf, n, n_limit, min_length)) { rec->status = SCAR_ST_FILTERED; continue; }
        const uint8_t *target = target_bytes + target_off[t];
        int L = (int)(target_off[t + 1] - target_off[t]);
        int cut = cutsite[t];
        /* cutwindow = target_region[cut-4:cut+4] (INDEL_Processing.py:172) */
        int cs, cl = py_slice(L, cut - 4, cut + 4, &cs);
        orc_sliding_window(buf, n, target, L, cut, 15, target + cs, cl, hr_donor, hlen, rec);
    }
    free(buf);
}
