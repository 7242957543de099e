/*
 * scar_synth.c — deterministic synthetic PEAR-merged amplicon reads (bench/test input generator).
 *
 * Not part of the product path and not derived from the reference: it only manufactures inputs of
 * the shapes BASELINE.json names.  Read r of a run is a pure function of
 * (seed, r, parameters): a counter-based splitmix64 stream per read, so any range of reads can be
 * generated independently (ranks generate their own shard; threads split a range).
 *
 *   read = target[cut-dl-a : cut-dl] + ins + target[cut+dr : cut+dr+b],  a + |ins| + b = len
 *   a = len/2 - phi - ceil(|ins|/2),  phi in 0..5 (primer phasing offset)
 */
#include <stdint.h>
#include <string.h>

typedef struct synth_params {
    uint64_t seed;
    int32_t read_len;       /* maximum (and usual) read length */
    int32_t min_len;        /* == read_len for fixed-length runs; else length uniform in [min_len, read_len] */
    int32_t n_targets;
    int32_t n_samples;      /* 0: no index fields */
    int32_t idx_len;        /* bytes per index field in the output (8 for Illumina dual indices) */
    int32_t mode;           /* 0 standard spectrum (C1/C2/C5), 1 microhomology/insertion heavy (C3), 2 large deletions (C4) */
    int32_t n_pool;         /* insertion pool size (256) */
    int32_t pool_stride;    /* bytes per pool entry (first byte = length) */
    double p_wt;            /* wild-type fraction */
    double p_n;             /* reads with one 'N' */
    double p_sub;           /* reads with one substitution near a junction */
    double p_idx1;          /* reads with one substitution in an index field */
    double p_idx2;          /* reads with two substitutions in one index field */
} synth_params;

static inline uint64_t splitmix(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t *s) { return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }
static inline int32_t below(uint64_t *s, int32_t n) { return (int32_t)(splitmix(s) % (uint64_t)n); }

static int32_t geometric(uint64_t *s, double mean, int32_t cap)
{
    /* number of failures before a success with p = 1/(mean+1); integer loop keeps it libm-free */
    double p = 1.0 / (mean + 1.0);
    int32_t k = 0;
    while (k < cap && u01(s) >= p) ++k;
    return k;
}

static const char ACGT[4] = {'A', 'C', 'G', 'T'};

/*
 * targets: CSR; cutsite per target; sample_target[n_samples] (or NULL: target = read % n_targets);
 * idx_a/idx_b: [n_samples][idx_len] index sequences written into the header fields;
 * mh_pairs: per target list of (x, y) int32 pairs with target[x]==target[y], x < cut <= y  (mode 1);
 *           mh_off[n_targets+1] indexes pairs.
 * outputs: seq [n][read_len] (padded with 0 past the read's length), len[n],
 *          f0/f1 [n][idx_len], truth_sample[n].
 */
void synth_reads(const synth_params *P, int64_t first, int64_t n,
                 const uint8_t *target_bytes, const int64_t *target_off, const int32_t *cutsite,
                 const int32_t *sample_target, const uint8_t *idx_a, const uint8_t *idx_b,
                 const uint8_t *pool, const int32_t *mh_pairs, const int64_t *mh_off,
                 uint8_t *seq, int32_t *len_out, uint8_t *f0, uint8_t *f1, int32_t *truth_sample)
{
    const int32_t RL = P->read_len;
    for (int64_t k = 0; k < n; ++k) {
        const int64_t r = first + k;
        uint64_t s = P->seed * 0xD1342543DE82EF95ull + (uint64_t)r * 0x2545F4914F6CDD1Dull + 0x1234567ull;
        (void)splitmix(&s);
        uint8_t *out = seq + k * (int64_t)RL;
        int32_t sample = -1, t;
        if (P->n_samples > 0) {
            sample = below(&s, P->n_samples);
            t = sample_target ? sample_target[sample] : sample % P->n_targets;
        } else {
            t = (int32_t)(r % P->n_targets);
        }
        if (truth_sample) truth_sample[k] = sample;
        const uint8_t *T = target_bytes + target_off[t];
        const int32_t L = (int32_t)(target_off[t + 1] - target_off[t]);
        const int32_t cut = cutsite[t];
        int32_t len = RL;
        if (P->min_len < RL) len = P->min_len + below(&s, RL - P->min_len + 1);

        int32_t dl = 0, dr = 0, il = 0;
        const uint8_t *ins = 0;
        const int wt = u01(&s) < P->p_wt;
        if (!wt) {
            if (P->mode == 1 && u01(&s) < 0.70 && mh_off && mh_off[t + 1] > mh_off[t]) {
                int64_t np = mh_off[t + 1] - mh_off[t];
                int64_t j = mh_off[t] + (int64_t)(splitmix(&s) % (uint64_t)np);
                int32_t x = mh_pairs[2 * j], y = mh_pairs[2 * j + 1];
                dl = cut - x; dr = y - cut;       /* read = T[..x) + T[y..) */
            } else if (P->mode == 2) {
                int32_t tot = 100 + below(&s, 181);
                dl = below(&s, tot + 1); dr = tot - dl;
            } else {
                dl = geometric(&s, 6.0, 60); dr = geometric(&s, 6.0, 60);
            }
            double u = u01(&s);
            double p_big = P->mode == 1 ? 0.25 : 0.10;
            if (u < p_big) {            /* 5-20 nt insertion from the upper half of the pool */
                int32_t e = P->n_pool / 2 + below(&s, P->n_pool - P->n_pool / 2);
                ins = pool + (int64_t)e * P->pool_stride + 1; il = pool[(int64_t)e * P->pool_stride];
            } else if (u < p_big + 0.15) { /* 1-3 nt insertion from the lower half */
                int32_t e = below(&s, P->n_pool / 2);
                ins = pool + (int64_t)e * P->pool_stride + 1; il = pool[(int64_t)e * P->pool_stride];
            }
        }
        const int32_t phi = below(&s, 6);
        int32_t a = len / 2 - phi - (il + 1) / 2;
        int32_t b = len - a - il;
        /* keep both flanks inside the locus; shrink the deletion if the locus is too short */
        while (dl > 0 && cut - dl - a < 0) --dl;
        while (dr > 0 && cut + dr + b > L) --dr;
        int32_t ls = cut - dl - a;
        if (ls < 0) { a += ls; b -= ls; ls = 0; }            /* degenerate tiny loci */
        if (cut + dr + b > L) b = L - cut - dr;
        int32_t w = 0;
        for (int32_t i = 0; i < a; ++i) out[w++] = T[ls + i];
        for (int32_t i = 0; i < il; ++i) out[w++] = ins[i];
        for (int32_t i = 0; i < b; ++i) out[w++] = T[cut + dr + i];
        while (w < len) out[w++] = (uint8_t)ACGT[below(&s, 4)];
        if (u01(&s) < P->p_sub) {       /* sequencing error within 10 nt of the junction */
            int32_t p = a - 10 + below(&s, 20 + il);
            if (p >= 0 && p < len) out[p] = (uint8_t)ACGT[below(&s, 4)];
        }
        if (u01(&s) < P->p_n) out[below(&s, len)] = 'N';
        for (int32_t i = len; i < RL; ++i) out[i] = 0;
        len_out[k] = len;

        if (P->n_samples > 0 && f0) {
            uint8_t *g0 = f0 + k * (int64_t)P->idx_len, *g1 = f1 + k * (int64_t)P->idx_len;
            memcpy(g0, idx_a + (int64_t)sample * P->idx_len, (size_t)P->idx_len);
            memcpy(g1, idx_b + (int64_t)sample * P->idx_len, (size_t)P->idx_len);
            double u = u01(&s);
            int nsub = u < P->p_idx2 ? 2 : (u < P->p_idx2 + P->p_idx1 ? 1 : 0);
            uint8_t *g = below(&s, 2) ? g1 : g0;
            for (int j = 0; j < nsub; ++j) {
                int32_t p = below(&s, P->idx_len);
                uint8_t c = (uint8_t)ACGT[below(&s, 4)];
                if (c == g[p]) c = (uint8_t)ACGT[(below(&s, 3) + 1 + (c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3)) & 3];
                g[p] = c;
            }
        }
    }
}
