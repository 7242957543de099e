// scar_api.cu — the C ABI of libscar_b200.so (include/scar_b200.h): context set-up (what
// window_mapping / dictionary_build / TargetMapper.phasing compute once per run, reference
// scarmapper/INDEL_Processing.py:57-80,1063-1124 and TargetMapper.py:30-71), the stage entry points
// and the host-buffer pipeline.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <string>
#include <atomic>
#include <thread>
#include <mutex>
#include <vector>

#include "scar_internal.cuh"

// ---- errors ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
int scar_fail(int code, const char *fmt, ...)      // for the other translation units
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
#include "scar_ctx.cuh"
#define fail scar_fail

extern "C" const char *scar_last_error(void) { return g_err; }
extern "C" int scar_version(void) { return 100; }

// ---- host helpers ---------------------------------------------------------------------------------
static uint8_t comp_host(uint8_t c)
{
    static const char *from = "ACGTMRWSYKVHDBXNacgtmrwsykvhdbxn", *to = "TGCAKYWSRMBDHVXNtgcakywsrmbdhvxn";
    const char *p = strchr(from, c);
    return (p && c) ? (uint8_t)to[p - from] : c;
}

extern "C" int scar_cutsite_search(const uint8_t *target, int32_t len, const uint8_t *sgrna, int32_t sg_len,
                                   int32_t reverse_comp, int32_t *cutsite)
{
    if (!target || !sgrna || !cutsite || len < 0 || sg_len < 0) return fail(SCAR_ERR_INVALID, "bad argument");
    std::string work(sg_len, 'N');
    for (int i = 0; i < sg_len; ++i) work[i] = reverse_comp ? (char)comp_host(sgrna[sg_len - 1 - i]) : (char)sgrna[i];
    // INDEL_Processing.py:773-791: windows [lft, lft+sg_len) while rt < len-1, first occurrence
    for (int lft = 0; lft + sg_len < len - 1; ++lft) {
        if (memcmp(target + lft, work.data(), (size_t)sg_len) == 0) {
            *cutsite = reverse_comp ? lft + 3 : lft + sg_len - 3;
            return SCAR_OK;
        }
    }
    return fail(SCAR_ERR_INVALID, "sgRNA does not map to the target region");
}

static bool acgt_only(const uint8_t *p, int n)
{
    for (int i = 0; i < n; ++i) if (!scar_base_valid(p[i])) return false;
    return true;
}
static uint64_t encode_bases(const uint8_t *p, int n)
{
    uint64_t c = 0;
    for (int i = 0; i < n; ++i) c |= (uint64_t)scar_base_code(p[i]) << (2 * i);
    return c;
}

// ---- "10-mer -> payload" tables of the static blob (layouts: ScarTableDesc, scar_internal.cuh) -------------------
static void append_aligned(std::vector<unsigned char> &bytes, const void *p, size_t n, uint32_t *off)
{
    bytes.resize((bytes.size() + 15) & ~(size_t)15);
    *off = (uint32_t)bytes.size();
    bytes.insert(bytes.end(), (const unsigned char *)p, (const unsigned char *)p + n);
}

// bucketed perfect hash (two entries per 8-byte bucket, one probe)
static bool build_bucketed(const std::vector<std::pair<uint32_t, uint32_t>> &kv /* key, payload */,
                           std::vector<unsigned char> &bytes, ScarTableDesc &out)
{
    // bucket = mulhi((key << 12) * mult, B): any B works, not only powers of two.  With two entries per bucket a
    // collision-free multiplier exists with fair probability once n^3 / (6 B^2) is a few units, so start at
    // B ~ n^1.5 / 5.5 (at least 1.5 n) and grow by 10 % per round of attempts.
    const size_t n = kv.size();
    double want = std::max(1.5 * (double)n, std::pow((double)n, 1.5) / 5.5);
    uint32_t B = std::max<uint32_t>(16, ((uint32_t)want + 1) & ~1u);
    uint64_t rng = 0x853C49E6748FEA9Bull ^ (n * 0x9E3779B97F4A7C15ull);
    for (int round = 0; round < 40 && B <= (1u << 17); ++round, B = ((uint32_t)(B * 1.1) + 2) & ~1u) {
        for (int attempt = 0; attempt < 3000; ++attempt) {
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            const uint32_t mult = (uint32_t)(rng >> 32) | 1u;
            std::vector<uint2> b(B, make_uint2(SCAR_TABLE_EMPTY, SCAR_TABLE_EMPTY));
            bool ok = true;
            for (const auto &e : kv) {
                const uint32_t t = e.first << 12;                  // key in the top 20 bits, as the kernel forms it
                const uint32_t slot = scar_mulhi32(t * mult, B);
                const uint32_t val = t | e.second;
                if (b[slot].x == SCAR_TABLE_EMPTY) b[slot].x = val;
                else if (b[slot].y == SCAR_TABLE_EMPTY) b[slot].y = val;
                else { ok = false; break; }
            }
            if (ok) {
                memset(&out, 0, sizeof(out));
                append_aligned(bytes, b.data(), b.size() * sizeof(uint2), &out.off);
                out.mult = mult; out.n = B;
                return true;
            }
        }
    }
    return false;
}

// compact hash-and-displace table: the keys of a displacement bucket are moved together (one displacement per
// bucket, largest buckets first) until they all fall on free slots
static bool build_compact(const std::vector<std::pair<uint32_t, uint32_t>> &kv, std::vector<unsigned char> &bytes, ScarTableDesc &out)
{
    const size_t n = kv.size();
    uint32_t N = std::max<uint32_t>(8, (uint32_t)((double)n / 0.72) + 1);
    const uint32_t nb = std::max<uint32_t>(4, (uint32_t)(n / 3) + 1);
    uint64_t rng = 0xD1B54A32D192ED03ull ^ (n * 0x9E3779B97F4A7C15ull);
    for (int round = 0; round < 24; ++round, N += N / 16 + 1) {
        const uint32_t unit = (uint32_t)((0x100000000ull + N - 1) / N);
        for (int attempt = 0; attempt < 16; ++attempt) {
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            const uint32_t mult = (uint32_t)(rng >> 32) | 1u;
            std::vector<std::vector<uint32_t>> members(nb);
            for (size_t i = 0; i < n; ++i) members[scar_chd_bucket((kv[i].first << 12) * mult, nb)].push_back((uint32_t)i);
            std::vector<uint32_t> order(nb);
            for (uint32_t i = 0; i < nb; ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return members[a].size() > members[b].size(); });
            std::vector<uint32_t> ent(N, SCAR_TABLE_EMPTY);
            std::vector<uint16_t> disp(nb, 0);
            bool ok = true;
            for (uint32_t bi = 0; bi < nb && ok; ++bi) {
                const auto &mem = members[order[bi]];
                if (mem.empty()) break;
                bool placed = false;
                for (uint32_t d = 0; d < 65536u && !placed; ++d) {
                    bool fit = true;
                    std::vector<uint32_t> slots;
                    for (uint32_t i : mem) {
                        const uint32_t sl = scar_chd_slot((kv[i].first << 12) * mult, d, unit, N);
                        if (ent[sl] != SCAR_TABLE_EMPTY || std::find(slots.begin(), slots.end(), sl) != slots.end()) { fit = false; break; }
                        slots.push_back(sl);
                    }
                    if (!fit) continue;
                    for (size_t q = 0; q < mem.size(); ++q) ent[slots[q]] = (kv[mem[q]].first << 12) | kv[mem[q]].second;
                    disp[order[bi]] = (uint16_t)d;
                    placed = true;
                }
                ok = placed;
            }
            if (!ok) continue;
            memset(&out, 0, sizeof(out));
            append_aligned(bytes, ent.data(), ent.size() * sizeof(uint32_t), &out.off);
            append_aligned(bytes, disp.data(), disp.size() * sizeof(uint16_t), &out.doff);
            out.mult = mult; out.n = N; out.nb = nb; out.unit = unit;
            return true;
        }
    }
    return false;
}

static bool build_table(const std::vector<std::pair<uint32_t, uint32_t>> &kv, bool compact, std::vector<unsigned char> &bytes,
                        ScarTableDesc &out)
{
    return compact ? build_compact(kv, bytes, out) : build_bucketed(kv, bytes, out);
}

// ---- context --------------------------------------------------------------------------------------
// (struct HostBuf / struct scar_ctx: scar_ctx.cuh)

static void free_hostbuf(HostBuf &b)
{
    cudaFree(b.seq); cudaFree(b.f0); cudaFree(b.f1); cudaFree(b.seq_off); cudaFree(b.f0_off); cudaFree(b.f1_off);
    cudaFree(b.seq_len); cudaFree(b.f0_len); cudaFree(b.f1_len);
    cudaFree(b.cells); cudaFree(b.lens); cudaFree(b.samples); cudaFree(b.records);
    if (b.stream) cudaStreamDestroy(b.stream);
    if (b.done) cudaEventDestroy(b.done);
    b = HostBuf();
}

extern "C" void scar_ctx_destroy(scar_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto &b : c->buf) free_hostbuf(b);
    cudaFree(c->d_sample_target); cudaFree(c->d_blob);
    cudaFree(c->d_class_map); cudaFree(c->d_right); cudaFree(c->d_left); cudaFree(c->d_peq); cudaFree(c->d_first_sample); cudaFree(c->d_fast_codes);
    cudaFree(c->d_lut_right); cudaFree(c->d_lut_left);
    cudaFree(c->d_slots); cudaFree(c->d_claimed); cudaFree(c->d_acc); cudaFree(c->d_status);
    cudaFree(c->d_fq_seq_off); cudaFree(c->d_fq_seq_len); cudaFree(c->d_gather_small); cudaFree(c->d_gather_out);
    for (uint8_t *p : c->h_piece) if (p) cudaFreeHost(p);
    cudaFree(c->d_text); cudaFree(c->d_tile_counts); cudaFree(c->d_tile_offsets); cudaFree(c->d_nl); cudaFree(c->d_fq_records); cudaFree(c->d_part_cursor);
    for (cudaEvent_t e : c->fq_events) cudaEventDestroy(e);
    delete c;
}

template <typename T> static int upload(T **dst, const std::vector<T> &v)
{
    const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    CU(cudaMalloc((void **)dst, bytes));
    if (!v.empty()) CU(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return SCAR_OK;
}

static int build_demux(scar_ctx *c, const scar_config *cfg)
{
    // distinct index strings per side, class map over every byte that occurs in an index string
    std::vector<uint8_t> cmap(256, 0);
    int n_classes = 1;
    auto str_of = [](const uint8_t *b, const int64_t *off, int s) {
        return std::string((const char *)b + off[s], (size_t)(off[s + 1] - off[s]));
    };
    std::vector<std::string> right_s, left_s;
    std::map<std::string, int> right_id, left_id;
    std::vector<int> s_right(c->n_samples), s_left(c->n_samples);
    for (int s = 0; s < c->n_samples; ++s) {
        std::string r = cfg->right_index_bytes ? str_of(cfg->right_index_bytes, cfg->right_index_off, s) : std::string();
        std::string l = cfg->left_index_bytes ? str_of(cfg->left_index_bytes, cfg->left_index_off, s) : std::string();
        if (r.size() > SCAR_MAX_INDEX_LEN || l.size() > SCAR_MAX_INDEX_LEN)
            return fail(SCAR_ERR_UNSUPPORTED, "index sequence longer than %d nt (sample %d)", SCAR_MAX_INDEX_LEN, s);
        if (!right_id.count(r)) { right_id[r] = (int)right_s.size(); right_s.push_back(r); }
        if (!left_id.count(l)) { left_id[l] = (int)left_s.size(); left_s.push_back(l); }
        s_right[s] = right_id[r]; s_left[s] = left_id[l];
        for (unsigned char ch : r + l)
            if (!cmap[ch]) {
                if (n_classes >= SCAR_MAX_CLASSES)
                    return fail(SCAR_ERR_UNSUPPORTED, "index sequences use more than %d distinct symbols", SCAR_MAX_CLASSES - 1);
                cmap[ch] = (uint8_t)n_classes++;
            }
    }
    if (right_s.size() > 64 || left_s.size() > 64)
        return fail(SCAR_ERR_UNSUPPORTED, "more than 64 distinct index sequences on one side (%zu, %zu)", right_s.size(), left_s.size());
    std::vector<uint64_t> peq;
    auto make = [&](const std::vector<std::string> &v, std::vector<DemuxString> &out) {
        for (const auto &s : v) {
            DemuxString d; d.len = (int)s.size(); d.peq_off = (int)peq.size(); d.code[0] = d.code[1] = 0;
            peq.resize(peq.size() + n_classes, 0ull);
            for (int i = 0; i < d.len; ++i) {
                const int cl = cmap[(unsigned char)s[i]];
                peq[d.peq_off + cl] |= 1ull << i;
                if (i < 32) d.code[i >> 4] |= (uint64_t)cl << (4 * (i & 15));
            }
            out.push_back(d);
        }
    };
    std::vector<DemuxString> hr, hl;
    make(right_s, hr); make(left_s, hl);
    std::vector<uint16_t> first(std::max<size_t>(hr.size() * hl.size(), 1), SCAR_SAMPLE_NONE);
    for (int s = c->n_samples - 1; s >= 0; --s) first[s_right[s] * hl.size() + s_left[s]] = (uint16_t)s;
    c->n_right = (int)hr.size(); c->n_left = (int)hl.size(); c->n_classes = n_classes;
    // fast path: every distinct string is ACGT-only and <= 16 nt -> 2-bit Hamming shortcut in K2
    std::vector<uint8_t> fast((hr.size() + hl.size()) * 16 + 1, 0);
    bool all_fast = true;
    for (size_t side = 0; side < 2; ++side) {
        const auto &v = side ? left_s : right_s;
        for (size_t d = 0; d < v.size(); ++d) {
            if (v[d].size() > 16 || !acgt_only((const uint8_t *)v[d].data(), (int)v[d].size())) { all_fast = false; continue; }
            for (size_t k = 0; k < v[d].size(); ++k)
                fast[(side ? hr.size() : 0) * 16 + d * 16 + k] = (uint8_t)scar_base_code((uint8_t)v[d][k]);
        }
    }
    c->demux_fast = all_fast ? 1 : 0;
    int rc;
    if ((rc = upload(&c->d_fast_codes, fast))) return rc;
    if ((rc = upload(&c->d_class_map, cmap))) return rc;
    if ((rc = upload(&c->d_right, hr))) return rc;
    if ((rc = upload(&c->d_left, hl))) return rc;
    if ((rc = upload(&c->d_peq, peq))) return rc;
    if ((rc = upload(&c->d_first_sample, first))) return rc;
    // Direct tables: when every distinct string of a side has the same length n <= 8 (Illumina 8, TruSeq 6),
    // the set of strings that accept a field is tabulated for every all-ACGT field of n bases (4^n entries,
    // indexed by the field's 2-bit code) with the arithmetic K2 uses (accepts(): match_maker <= mismatch).
    if (all_fast && c->platform != SCAR_PLATFORM_RAMSDEN) {
        for (int side = 0; side < 2; ++side) {
            const auto &v = side ? left_s : right_s;
            const auto &hs = side ? hl : hr;
            if (v.empty()) continue;
            const size_t n = v[0].size();
            bool same = n >= 1 && n <= 8;
            for (const auto &x : v) same = same && x.size() == n;
            if (!same) continue;
            std::vector<uint64_t> lut((size_t)1 << (2 * n));
            uint8_t u[8];
            const uint64_t no_code[2] = {0, 0};
            for (size_t code = 0; code < lut.size(); ++code) {
                for (size_t k = 0; k < n; ++k) u[k] = (uint8_t)"ACTG"[(code >> (2 * k)) & 3];
                uint64_t m = 0;
                for (size_t d = 0; d < v.size(); ++d) {
                    int h = 0;                         // equal lengths: Hamming decides for mismatch <= 1
                    for (size_t k = 0; k < n; ++k) h += u[k] != (uint8_t)v[d][k];
                    bool ok = h <= c->mismatch;
                    if (!ok && c->mismatch > 1)
                        ok = accepts(hs[d], peq.data(), u, (int)n, no_code, cmap.data(), c->mismatch, false);
                    if (ok) m |= 1ull << d;
                }
                lut[code] = m;
            }
            if ((rc = upload(side ? &c->d_lut_left : &c->d_lut_right, lut))) return rc;
            (side ? c->lut_n_left : c->lut_n_right) = (int)n;
        }
    }
    return SCAR_OK;
}

extern "C" int scar_ctx_create(const scar_config *cfg, scar_ctx **out)
{
    if (!cfg || !out) return fail(SCAR_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->n_targets < 1 || cfg->n_samples < 1 || cfg->n_samples >= 0xFFFF)
        return fail(SCAR_ERR_INVALID, "n_targets/n_samples out of range");
    if (cfg->platform < 0 || cfg->platform > 2) return fail(SCAR_ERR_INVALID, "unknown platform %d", cfg->platform);
    if (!cfg->target_bytes || !cfg->target_off || !cfg->target_cutsite || !cfg->sample_target)
        return fail(SCAR_ERR_INVALID, "target_bytes / target_off / target_cutsite / sample_target must not be NULL");
    if ((cfg->left_index_bytes && !cfg->left_index_off) || (cfg->right_index_bytes && !cfg->right_index_off))
        return fail(SCAR_ERR_INVALID, "index bytes given without offsets");
    if (cfg->phase_off && (!cfg->phase_r1_bytes || !cfg->phase_r1_off || !cfg->phase_r2_bytes || !cfg->phase_r2_off))
        return fail(SCAR_ERR_INVALID, "phase_off given without the phase string lists");
    if (cfg->hr_len > 0 && !cfg->hr_donor) return fail(SCAR_ERR_INVALID, "hr_len > 0 with a NULL donor");
    if (cfg->max_read_len < 1 || cfg->max_read_len > SCAR_MAX_READ_LEN)
        return fail(SCAR_ERR_UNSUPPORTED, "max_read_len %d outside 1..%d", cfg->max_read_len, SCAR_MAX_READ_LEN);
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(SCAR_ERR_CUDA, "CUDA device %d not present (%d devices)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail(SCAR_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);

    scar_ctx *c = new scar_ctx();
    struct Guard { scar_ctx *c; ~Guard() { if (c) scar_ctx_destroy(c); } } guard{c};
    c->device = cfg->device; c->sm_count = c->sm_total = prop.multiProcessorCount; c->platform = cfg->platform;
    c->mismatch = cfg->mismatch >= 0 ? cfg->mismatch : (cfg->platform == 0 ? 1 : cfg->platform == 1 ? 0 : 3);
    c->n_targets = cfg->n_targets; c->n_samples = cfg->n_samples;
    c->max_read_len = cfg->max_read_len; c->W = (cfg->max_read_len + 15) / 16;
    c->n_limit = cfg->n_limit; c->min_length = cfg->min_length;
    if (const char *e = getenv("SCAR_CHUNK_READS")) { long v = atol(e); if (v >= 1024) c->chunk_reads = v; }

    // HR donor (SlidingWindow.pyx:102-117)
    if (cfg->hr_len > 0) {
        if (cfg->hr_len > 32 || !acgt_only(cfg->hr_donor, cfg->hr_len))
            return fail(SCAR_ERR_UNSUPPORTED, "HR donor must be 1..32 nt of ACGT");
        c->hr_len = cfg->hr_len; c->hr_code = encode_bases(cfg->hr_donor, cfg->hr_len);
    }

    // targets: window tables (window_mapping, INDEL_Processing.py:57-80)
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> tab_pos(cfg->n_targets), tab_left(cfg->n_targets), tab_right(cfg->n_targets);
    std::vector<ScarPhase> phases;
    std::vector<std::vector<uint32_t>> locus_words;
    std::vector<std::vector<uint8_t>> phase_luts;
    for (int t = 0; t < cfg->n_targets; ++t) {
        const uint8_t *T = cfg->target_bytes + cfg->target_off[t];
        const int L = (int)(cfg->target_off[t + 1] - cfg->target_off[t]);
        const int cut = cfg->target_cutsite[t];
        if (cut < 0 || cut > L || L > 0xFFFF) return fail(SCAR_ERR_INVALID, "target %d: cutsite %d outside 0..%d", t, cut, L);
        ScarTargetMeta m; memset(&m, 0, sizeof(m));
        m.cutsite = cut; m.length = L;
        m.n_left = cut > 15 ? cut - 15 : 0;                 // while rt_position > 15
        m.n_right = L - 15 - cut > 0 ? L - 15 - cut : 0;    // while lft_position < L - 15
        if (m.n_left > SCAR_MAX_WINDOWS || m.n_right > SCAR_MAX_WINDOWS)
            return fail(SCAR_ERR_UNSUPPORTED, "target %d: more than %d windows on one side", t, SCAR_MAX_WINDOWS);
        // anchor-and-extend path: the locus is all ACGT, every 10-mer of it occurs once, both window lists
        // are non-empty and positions fit the 12-bit table payload
        bool fast = m.n_left >= 1 && m.n_right >= 1 && L >= 10 && L - 10 <= 4094 && acgt_only(T, L) &&
                    !getenv("SCAR_NO_ANCHOR");
        std::vector<std::pair<uint32_t, uint32_t>> pos_kv;
        if (fast) {
            std::map<uint32_t, uint32_t> seen;
            for (int p = 0; p + 10 <= L && fast; ++p) {
                const uint32_t key = (uint32_t)encode_bases(T + p, 10);
                if (seen.count(key)) fast = false; else seen[key] = (uint32_t)p;
            }
            if (fast) pos_kv.assign(seen.begin(), seen.end());
        }
        m.fast = fast ? 1 : 0;
        if (fast) {
            tab_pos[t] = pos_kv;
            // packed locus: one word (16 bases) of padding in front, eight behind (a 64-base compare step may
            // start up to 10 bases past the cut and reads three words per 32 bases from an aligned start)
            std::vector<uint32_t> tw((size_t)(L + 15) / 16 + 9, 0u);
            for (int p = 0; p < L; ++p) tw[1 + p / 16] |= scar_base_code(T[p]) << (2 * (p % 16));
            locus_words.push_back(tw);
        } else {
            locus_words.push_back(std::vector<uint32_t>());
        }
        for (int side = 0; side < 2 && !fast; ++side) {
            std::map<uint32_t, uint32_t> mn;
            const int n = side == 0 ? m.n_left : m.n_right;
            for (int i = 0; i < n; ++i) {
                const int a = side == 0 ? cut - 10 - i : cut + i;
                if (a < 0 || a + 10 > L) continue;          // Python slice shorter than 10: can never equal a read window
                if (!acgt_only(T + a, 10))
                    return fail(SCAR_ERR_UNSUPPORTED, "target %d: window at %d contains a non-ACGT base", t, a);
                const uint32_t key = (uint32_t)encode_bases(T + a, 10);
                if (!mn.count(key)) mn[key] = (uint32_t)i;  // list order: the smallest i wins
            }
            (side == 0 ? tab_left[t] : tab_right[t]).assign(mn.begin(), mn.end());
        }
        m.cutwindow_key = 0xFFFFFFFFu;
        if (cfg->cutwindow_bytes && acgt_only(cfg->cutwindow_bytes + 10 * t, 10))
            m.cutwindow_key = (uint32_t)encode_bases(cfg->cutwindow_bytes + 10 * t, 10);
        // phasing (TargetMapper.py:54-69 lists; compared at INDEL_Processing.py:965-970)
        m.phase_off = (int)phases.size(); m.n_phase = 0;
        if (cfg->phase_off && cfg->platform == SCAR_PLATFORM_ILLUMINA) {
            const int k0 = cfg->phase_off[t], k1 = cfg->phase_off[t + 1];
            if (k1 - k0 > SCAR_MAX_PHASES) return fail(SCAR_ERR_UNSUPPORTED, "target %d: more than %d phases", t, SCAR_MAX_PHASES);
            for (int k = k0; k < k1; ++k) {
                ScarPhase p; memset(&p, 0, sizeof(p));
                const uint8_t *a = cfg->phase_r1_bytes + cfg->phase_r1_off[k];
                const int na = (int)(cfg->phase_r1_off[k + 1] - cfg->phase_r1_off[k]);
                const uint8_t *b = cfg->phase_r2_bytes + cfg->phase_r2_off[k];
                const int nb = (int)(cfg->phase_r2_off[k + 1] - cfg->phase_r2_off[k]);
                if (na > 16 || nb > 16) return fail(SCAR_ERR_UNSUPPORTED, "phasing string longer than 16 nt");
                if (na == 0) p.r1_len = 0;
                else if (!acgt_only(a, na)) p.r1_len = 0xFF;
                else { p.r1_len = (uint8_t)na; p.r1_code = (uint32_t)encode_bases(a, na); }
                if (nb == 0) p.r2_len = 0;
                else if (!acgt_only(b, nb)) p.r2_len = 0xFF;
                else {
                    uint8_t rcb[16];
                    for (int i = 0; i < nb; ++i) rcb[i] = comp_host(b[nb - 1 - i]);   // read suffix == rcomp(R2 string)
                    p.r2_len = (uint8_t)nb; p.r2_code = (uint32_t)encode_bases(rcb, nb);
                }
                phases.push_back(p);
            }
            m.n_phase = k1 - k0;
            // phasing lookup tables when every non-empty string of a side has one length <= 6 and is ACGT
            int n1 = 0, n2 = 0; bool lut_ok = true;
            for (int k = k0; k < k1 && lut_ok; ++k) {
                const ScarPhase &p = phases[m.phase_off + (k - k0)];
                if (p.r1_len == 0) continue;                 // skipped phase (both sides)
                if (p.r1_len > 6 || (n1 && p.r1_len != n1)) lut_ok = false; else n1 = p.r1_len;
                if (p.r2_len != 0) { if (p.r2_len > 6 || (n2 && p.r2_len != n2)) lut_ok = false; else n2 = p.r2_len; }
            }
            if (lut_ok && n1) {
                const int stride = 1 << (2 * std::max(n1, n2));      // 4^n entries per side
                std::vector<uint8_t> lut((size_t)2 * stride, 0);
                for (int k = k1 - 1; k >= k0; --k) {          // descending so that the first k wins
                    const ScarPhase &p = phases[m.phase_off + (k - k0)];
                    if (p.r1_len == 0) continue;
                    lut[p.r1_code] = (uint8_t)(k - k0 + 1);
                    if (p.r2_len != 0) lut[stride + p.r2_code] = (uint8_t)(k - k0 + 1);
                }
                phase_luts.push_back(lut); m.ph_n1 = n1; m.ph_n2 = n2; m.ph_lut_off = 1;   // offset fixed up below
            } else phase_luts.push_back(std::vector<uint8_t>());
        } else phase_luts.push_back(std::vector<uint8_t>());
        if ((int)phase_luts.size() < t + 1) phase_luts.push_back(std::vector<uint8_t>());
        c->h_targets.push_back(m);
    }
    c->do_phasing = (cfg->phase_off && cfg->platform == SCAR_PLATFORM_ILLUMINA) ? 1 : 0;
    // one blob, staged into shared memory by a single bulk copy: tables | target metas | phases | nmax
    // nmax[len] = largest n with double(n)/double(len) <= N_Limit: the reference's filter
    // seq.count("N")/len(seq) > N_Limit (INDEL_Processing.py:147) evaluated once per length on the host
    // with the same IEEE-754 double division Python performs.
    // (tabulated for every supported length, so that a context can be re-planned for longer reads: scar_grow_read_len)
    std::vector<uint16_t> nmax((size_t)SCAR_MAX_READ_LEN + 1, 0);
    for (int len = 1; len <= SCAR_MAX_READ_LEN; ++len) {
        int n = 0;
        while (n < len && !((double)(n + 1) / (double)len > cfg->n_limit)) ++n;
        nmax[len] = (uint16_t)n;
        // monotone in n, so the first n+1 that exceeds the limit ends the search
    }
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    // The tables: bucketed (one probe) while they are small; a dozen loci would leave a CTA no room for more than
    // one tile beside them, so past 48 KB (or with SCAR_COMPACT_TABLES=1) the compact layout is built instead.
    std::vector<unsigned char> bytes;
    auto build_tables = [&](bool compact) -> int {
        bytes.clear();
        for (int t = 0; t < cfg->n_targets; ++t) {
            ScarTargetMeta &m = c->h_targets[t];
            if (m.fast) {
                if (!build_table(tab_pos[t], compact, bytes, m.pos)) return t + 1;
            } else if (!build_table(tab_left[t], compact, bytes, m.left) || !build_table(tab_right[t], compact, bytes, m.right)) return t + 1;
        }
        bytes.resize(align16(bytes.size()));
        return 0;
    };
    const char *force = getenv("SCAR_COMPACT_TABLES");
    c->compact_tables = force && atoi(force) != 0;
    int bad = build_tables(c->compact_tables);
    if (!bad && !c->compact_tables && !(force && atoi(force) == 0) && bytes.size() > 48 * 1024) {
        c->compact_tables = true;
        bad = build_tables(true);
    }
    if (bad) return fail(SCAR_ERR_UNSUPPORTED, "target %d: could not build a window table", bad - 1);
    for (int t = 0; t < cfg->n_targets; ++t) {
        if (!c->h_targets[t].fast) continue;
        c->h_targets[t].tw_off = (uint32_t)bytes.size();
        const auto &tw = locus_words[t];
        bytes.resize(align16(bytes.size() + tw.size() * sizeof(uint32_t)));
        memcpy(bytes.data() + c->h_targets[t].tw_off, tw.data(), tw.size() * sizeof(uint32_t));
    }
    for (int t = 0; t < cfg->n_targets; ++t) {
        if (!c->h_targets[t].ph_lut_off) continue;
        c->h_targets[t].ph_lut_off = (uint32_t)bytes.size();
        bytes.resize(align16(bytes.size() + phase_luts[t].size()));
        memcpy(bytes.data() + c->h_targets[t].ph_lut_off, phase_luts[t].data(), phase_luts[t].size());
    }
    c->meta_off = (uint32_t)bytes.size();
    bytes.resize(align16(bytes.size() + c->h_targets.size() * sizeof(ScarTargetMeta)));
    memcpy(bytes.data() + c->meta_off, c->h_targets.data(), c->h_targets.size() * sizeof(ScarTargetMeta));
    c->phase_off = (uint32_t)bytes.size();
    bytes.resize(align16(bytes.size() + std::max<size_t>(phases.size(), 1) * sizeof(ScarPhase)));
    if (!phases.empty()) memcpy(bytes.data() + c->phase_off, phases.data(), phases.size() * sizeof(ScarPhase));
    c->nmax_off = (uint32_t)bytes.size();
    bytes.resize(align16(bytes.size() + nmax.size() * sizeof(uint16_t)));
    memcpy(bytes.data() + c->nmax_off, nmax.data(), nmax.size() * sizeof(uint16_t));
    c->blob_bytes = (uint32_t)bytes.size();
    // larger blobs (many loci) are read through the read-only global path instead of shared memory
    c->smem_tables = c->blob_bytes <= 200 * 1024 && !getenv("SCAR_FORCE_GLOBAL_TABLES");

    int rc;
    if ((rc = upload(&c->d_blob, bytes))) return rc;
    std::vector<int32_t> st(cfg->sample_target, cfg->sample_target + cfg->n_samples);
    for (int s = 0; s < cfg->n_samples; ++s)
        if (st[s] < 0 || st[s] >= cfg->n_targets) return fail(SCAR_ERR_INVALID, "sample %d: target id %d out of range", s, st[s]);
    if ((rc = upload(&c->d_sample_target, st))) return rc;
    if ((rc = build_demux(c, cfg))) return rc;

    // accumulators
    uint64_t cap = cfg->table_capacity > 0 ? (uint64_t)cfg->table_capacity : (1ull << 22);
    if (cap & (cap - 1)) return fail(SCAR_ERR_INVALID, "table_capacity must be a power of two");
    c->capacity = cap;
    CU(cudaMalloc((void **)&c->d_slots, cap * sizeof(ScarTableSlot)));
    CU(cudaMalloc((void **)&c->d_claimed, cap * sizeof(uint32_t)));
    c->acc_words = 8 + (size_t)c->n_samples * SCAR_N_COUNTERS + (size_t)c->n_samples * 2 * SCAR_PHASE_BINS;
    CU(cudaMalloc((void **)&c->d_acc, c->acc_words * sizeof(unsigned long long)));
    c->d_n_entries = c->d_acc; c->d_unassigned = c->d_acc + 1; c->d_collisions = c->d_acc + 2; c->d_export_n = c->d_acc + 3;
    c->d_counters = c->d_acc + 8; c->d_phase_hist = c->d_counters + (size_t)c->n_samples * SCAR_N_COUNTERS;
    CU(cudaMalloc((void **)&c->d_status, sizeof(int32_t)));
    if ((rc = scar_reset(c))) return rc;
    guard.c = nullptr;
    *out = c;
    return SCAR_OK;
}

extern "C" int scar_reset(scar_ctx *c)
{
    if (!c) return fail(SCAR_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaMemset(c->d_slots, 0, c->capacity * sizeof(ScarTableSlot)));
    CU(cudaMemset(c->d_acc, 0, c->acc_words * sizeof(unsigned long long)));
    CU(cudaMemset(c->d_status, 0, sizeof(int32_t)));
    CU(cudaDeviceSynchronize());
    return SCAR_OK;
}

extern "C" int scar_reset_dev(scar_ctx *c, void *stream)
{
    if (!c) return fail(SCAR_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    // only the claimed slots are emptied (the list is complete: every insert path appends to it)
    CU(scar_launch_clear(c->d_slots, c->d_claimed, c->d_n_entries, c->sm_count, s));
    ++c->launches;
    CU(cudaMemsetAsync(c->d_acc, 0, c->acc_words * sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(c->d_status, 0, sizeof(int32_t), s));      // like scar_reset: a reset state carries no sticky error
    return SCAR_OK;
}

static K4Params k4_params(scar_ctx *c);

// Re-plan the context for reads up to max_len (the FASTQ file path learns the longest read while it streams).
// Nothing that has been computed depends on W: records, tables and counters stay valid.  The caller's streams are idle.
int scar_grow_read_len(scar_ctx *c, int max_len)
{
    if (max_len <= c->max_read_len) return SCAR_OK;
    if (max_len > SCAR_MAX_READ_LEN) return fail(SCAR_ERR_UNSUPPORTED, "a read of %d nt is longer than the supported %d", max_len, SCAR_MAX_READ_LEN);
    c->max_read_len = std::min(SCAR_MAX_READ_LEN, (max_len + 15) / 16 * 16);
    c->W = (c->max_read_len + 15) / 16;
    c->fused_ok = -1;
    for (auto &b : c->buf) {               // the packed batch of the three-kernel path is W cells per read
        cudaFree(b.cells); b.cells = nullptr;
        if (b.reads_cap > 0 && !scar_fused_path(c)) CU(cudaMalloc((void **)&b.cells, (size_t)b.reads_cap * c->W * sizeof(uint64_t)));
    }
    return SCAR_OK;
}

// Room for n_more further scar patterns: when the table would pass half full it is rebuilt at a larger power of two
// (entries exported through the claimed-slot list, folded into the new slots).  Synchronises the device.
extern "C" int scar_table_reserve(scar_ctx *c, int64_t n_more)
{
    if (!c || n_more < 0) return fail(SCAR_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    unsigned long long n = 0;
    CU(cudaMemcpy(&n, c->d_n_entries, sizeof(n), cudaMemcpyDeviceToHost));
    const uint64_t want = 2 * (n + (uint64_t)n_more);
    if (want <= c->capacity) return SCAR_OK;
    uint64_t cap = c->capacity;
    while (cap < 2 * want) cap <<= 1;      // ends at most a quarter full
    if (cap > (1ull << 32)) { cap = 1ull << 32; if (want > cap) return fail(SCAR_ERR_TABLE_FULL, "scar table cannot grow past 2^32 slots"); }
    scar_entry *d = nullptr;
    if (n) {
        CU(cudaMalloc((void **)&d, (size_t)n * sizeof(scar_entry)));
        CU(scar_launch_export(c->d_slots, c->d_claimed, c->d_n_entries, d, n, c->sm_count, nullptr));
        CU(cudaDeviceSynchronize());
    }
    cudaFree(c->d_slots); cudaFree(c->d_claimed);
    c->d_slots = nullptr; c->d_claimed = nullptr;
    cudaError_t e = cudaMalloc((void **)&c->d_slots, cap * sizeof(ScarTableSlot));
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_claimed, cap * sizeof(uint32_t));
    if (e != cudaSuccess) { cudaFree(d); return fail(SCAR_ERR_NOMEM, "scar table: no device memory for %llu slots", (unsigned long long)cap); }
    c->capacity = cap;
    CU(cudaMemset(c->d_slots, 0, cap * sizeof(ScarTableSlot)));
    CU(cudaMemset(c->d_n_entries, 0, sizeof(unsigned long long)));
    if (n) {
        CU(scar_launch_merge(k4_params(c), d, n, c->sm_count, nullptr));
        ++c->launches;
        CU(cudaDeviceSynchronize());
        cudaFree(d);
    }
    return scar_check_status(c);
}

extern "C" int scar_ctx_set_sm_limit(scar_ctx *c, int32_t n_sms)
{
    if (!c) return fail(SCAR_ERR_INVALID, "null context");
    if (n_sms <= 0 || n_sms > c->sm_total) n_sms = c->sm_total;
    c->sm_count = n_sms;
    return SCAR_OK;
}

extern "C" int scar_ctx_set_verify(scar_ctx *c, int32_t on)
{
    if (!c) return fail(SCAR_ERR_INVALID, "null context");
    c->verify = on != 0;
    c->weak_keys = on == 2;                 // test hook: tags that ignore the insertion / microhomology bytes collide
    return SCAR_OK;
}

extern "C" int scar_device_memory(int32_t device, int64_t *free_bytes, int64_t *total_bytes)
{
    CU(cudaSetDevice(device));
    size_t f = 0, t = 0;
    CU(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    return SCAR_OK;
}

extern "C" int64_t scar_launch_count(const scar_ctx *c) { return c ? c->launches : 0; }

extern "C" int scar_cells_per_read(const scar_ctx *c) { return c ? c->W : 0; }

extern "C" int scar_ctx_info(const scar_ctx *c, int64_t *static_bytes, int32_t *tables_in_smem, int32_t *n_anchor_targets)
{
    if (!c) return fail(SCAR_ERR_INVALID, "null context");
    if (static_bytes) *static_bytes = c->blob_bytes;
    if (tables_in_smem) *tables_in_smem = c->smem_tables ? 1 : 0;
    if (n_anchor_targets) { int k = 0; for (const auto &m : c->h_targets) k += m.fast; *n_anchor_targets = k; }
    return SCAR_OK;
}

extern "C" int scar_window_counts(const scar_ctx *c, int32_t t, int32_t *n_left, int32_t *n_right)
{
    if (!c || t < 0 || t >= c->n_targets) return fail(SCAR_ERR_INVALID, "bad target");
    if (n_left) *n_left = c->h_targets[t].n_left;
    if (n_right) *n_right = c->h_targets[t].n_right;
    return SCAR_OK;
}

// ---- stage entry points (device pointers) ---------------------------------------------------------
extern "C" int scar_pack_dev(scar_ctx *c, const scar_ragged *seq, int64_t n, uint64_t *cells, uint32_t *lens, void *stream)
{
    if (!c || !seq || n < 0) return fail(SCAR_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    CU(scar_launch_pack(*seq, n, c->W, c->platform, cells, n, lens, c->d_status, c->sm_count, (cudaStream_t)stream, &c->launches));
    return SCAR_OK;
}

extern "C" int scar_demux_dev(scar_ctx *c, const scar_ragged *seq, const scar_ragged *f0, const scar_ragged *f1,
                              int64_t n, uint16_t *samples, void *stream)
{
    if (!c || n < 0 || !samples) return fail(SCAR_ERR_INVALID, "bad argument");
    if (c->platform == SCAR_PLATFORM_ILLUMINA ? (!f0 || !f1) : !seq)
        return fail(SCAR_ERR_INVALID, "demux inputs missing for platform %d", c->platform);
    CU(cudaSetDevice(c->device));
    K2Params P; memset(&P, 0, sizeof(P));
    if (seq) P.seq = *seq; if (f0) P.f0 = *f0; if (f1) P.f1 = *f1;
    P.samples = samples; P.n_reads = n; P.class_map = c->d_class_map; P.right = c->d_right; P.left = c->d_left;
    P.peq = c->d_peq; P.first_sample = c->d_first_sample; P.n_right = c->n_right; P.n_left = c->n_left;
    P.n_classes = c->n_classes; P.platform = c->platform; P.mismatch = c->mismatch;
    P.fast_codes = c->d_fast_codes; P.fast = c->demux_fast;
    P.lut_right = c->d_lut_right; P.lut_left = c->d_lut_left; P.lut_n_right = c->lut_n_right; P.lut_n_left = c->lut_n_left;
    CU(scar_launch_demux(P, c->sm_count, (cudaStream_t)stream, &c->launches));
    return SCAR_OK;
}

static K3Params k3_params(const scar_ctx *c)
{
    K3Params P; memset(&P, 0, sizeof(P));
    P.sample_target = c->d_sample_target;
    P.blob = c->d_blob; P.blob_bytes = c->blob_bytes; P.meta_off = c->meta_off; P.phase_off = c->phase_off;
    P.nmax_off = c->nmax_off; P.W = c->W; P.n_samples = c->n_samples;
    P.do_phasing = c->do_phasing; P.min_length = c->min_length; P.hr_code = c->hr_code; P.hr_len = c->hr_len;
    P.compact_tables = c->compact_tables ? 1 : 0;
    return P;
}

extern "C" int scar_k3_plan(const scar_ctx *c, int32_t *cells_staged, int32_t *ctas_per_sm, int32_t *smem_bytes)
{
    if (!c) return fail(SCAR_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->device));
    int st = 0, per = 0, sm = 0;
    CU(scar_window_plan(k3_params(c), c->smem_tables, &st, &per, &sm));
    if (cells_staged) *cells_staged = st;
    if (ctas_per_sm) *ctas_per_sm = per;
    if (smem_bytes) *smem_bytes = sm;
    return SCAR_OK;
}

extern "C" int scar_classify_dev(scar_ctx *c, const uint64_t *cells, const uint32_t *lens, const uint16_t *samples,
                                 int64_t n, scar_record *records, void *stream)
{
    if (!c || n < 0) return fail(SCAR_ERR_INVALID, "bad argument");
    if (n == 0) return SCAR_OK;
    CU(cudaSetDevice(c->device));
    K3Params P = k3_params(c);
    P.cells = cells; P.lens = lens; P.samples = samples; P.records = records; P.n_reads = n; P.stride = n;
    CU(scar_launch_window(P, c->smem_tables, c->sm_count, (cudaStream_t)stream, &c->launches));
    return SCAR_OK;
}

static K4Params k4_params(scar_ctx *c)
{
    K4Params P; memset(&P, 0, sizeof(P));
    P.slots = c->d_slots; P.slot_mask = c->capacity - 1; P.claimed = c->d_claimed; P.n_entries = c->d_n_entries; P.counters = c->d_counters;
    P.unassigned = c->d_unassigned; P.phase_hist = c->d_phase_hist; P.status = c->d_status; P.n_samples = c->n_samples;
    P.platform = c->platform; P.max_probe = (uint32_t)std::min<uint64_t>(c->capacity - 1, 1u << 16);
    int max_phase = 0;
    for (const auto &m : c->h_targets) max_phase = std::max(max_phase, (int)m.n_phase);
    P.weak_keys = c->weak_keys ? 1 : 0;
    P.phase_bins = max_phase + 2;                 // bins: none, 0..max_phase-1, n/a
    P.phase_joint = P.phase_bins * P.phase_bins <= 128 ? 1 : 0;
    return P;
}

extern "C" int scar_aggregate_dev(scar_ctx *c, const scar_ragged *seq, const scar_record *records, int64_t n,
                                  uint64_t first_index, void *stream)
{
    if (!c || !seq || n < 0) return fail(SCAR_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    K4Params P = k4_params(c);
    P.seq = *seq; P.records = records; P.n_reads = n; P.first_index = first_index;
    CU(scar_launch_aggregate(P, c->sm_count, (cudaStream_t)stream, &c->launches));
    return SCAR_OK;
}

// parameters of the fused kernel (K1 + K3 and, with `aggregate`, K4 in one pass over the raw bytes)
static K3Params fused_params(scar_ctx *c, const scar_ragged *seq, const uint16_t *samples, int64_t n, uint64_t first_index,
                             scar_record *records, int aggregate)
{
    K3Params P = k3_params(c);
    if (seq) P.seq = *seq;
    P.samples = samples; P.records = records; P.n_reads = n; P.stride = n; P.platform = c->platform;
    P.aggregate = aggregate | (c->weak_keys ? 2 : 0); P.first_index = first_index;
    const K4Params A = k4_params(c);
    P.slots = A.slots; P.slot_mask = A.slot_mask; P.max_probe = A.max_probe; P.claimed = A.claimed; P.n_entries = A.n_entries;
    P.counters = A.counters; P.unassigned = A.unassigned; P.phase_hist = A.phase_hist; P.status = c->d_status;
    P.phase_bins = A.phase_bins; P.phase_joint = A.phase_joint;
    return P;
}

// does scar_run_dev take the fused kernel for this context?  (eligible platform / read length, and a tile fits beside
// the static blob in shared memory; independent of the layout of the reads)
bool scar_fused_path(scar_ctx *c)
{
    if (c->fused_ok < 0) {
        c->fused_ok = 0;
        if (scar_fused_eligible(c->W, c->platform)) {
            cudaSetDevice(c->device);
            int per = 0, sm = 0;
            if (scar_fused_plan(fused_params(c, nullptr, nullptr, 0, 0, nullptr, 1), c->smem_tables, c->max_read_len, &per, &sm) == cudaSuccess)
                c->fused_ok = 1;
            else cudaGetLastError();
        }
    }
    return c->fused_ok == 1;
}

extern "C" int scar_fused_info(const scar_ctx *c, int32_t *eligible, int32_t *ctas_per_sm, int32_t *smem_bytes)
{
    if (!c) return fail(SCAR_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->device));
    const bool ok = scar_fused_path(const_cast<scar_ctx *>(c));
    if (eligible) *eligible = ok ? 1 : 0;
    int per = 0, sm = 0;
    if (ok) CU(scar_fused_plan(fused_params(const_cast<scar_ctx *>(c), nullptr, nullptr, 0, 0, nullptr, 1), c->smem_tables,
                               c->max_read_len, &per, &sm));
    if (ctas_per_sm) *ctas_per_sm = per;
    if (smem_bytes) *smem_bytes = sm;
    return SCAR_OK;
}

extern "C" int scar_classify_raw_dev(scar_ctx *c, const scar_ragged *seq, const uint16_t *samples, int64_t n,
                                     uint64_t first_index, scar_record *records, int32_t aggregate, void *stream)
{
    if (!c || !seq || !samples || !records || n < 0) return fail(SCAR_ERR_INVALID, "bad argument");
    if (!scar_fused_eligible(c->W, c->platform))
        return fail(SCAR_ERR_UNSUPPORTED, "the fused kernel takes reads up to 320 nt on the Illumina / TruSeq platforms "
                                          "(use scar_pack_dev + scar_classify_dev + scar_aggregate_dev)");
    if (n == 0) return SCAR_OK;
    CU(cudaSetDevice(c->device));
    CU(scar_launch_fused(fused_params(c, seq, samples, n, first_index, records, aggregate ? 1 : 0), c->smem_tables,
                         c->max_read_len, c->sm_count, (cudaStream_t)stream, &c->launches));
    return SCAR_OK;
}

extern "C" int scar_run_dev(scar_ctx *c, const scar_ragged *seq, const scar_ragged *f0, const scar_ragged *f1, int64_t n,
                            uint64_t first_index, uint64_t *cells, uint32_t *lens, uint16_t *samples,
                            scar_record *records, void *stream)
{
    int rc;
    if ((rc = scar_demux_dev(c, seq, f0, f1, n, samples, stream))) return rc;
    // one pass over the raw bytes (pack, classify, aggregate) when the fused kernel takes the problem
    // (ineligible, or a static blob too large to leave room for a tile beside it: K1 -> K3 -> K4 below)
    if (c && seq && scar_fused_path(c)) return scar_classify_raw_dev(c, seq, samples, n, first_index, records, 1, stream);
    if (!cells || !lens) return fail(SCAR_ERR_INVALID, "the three-kernel path needs the cells / lens scratch buffers");
    if ((rc = scar_pack_dev(c, seq, n, cells, lens, stream))) return rc;
    if ((rc = scar_classify_dev(c, cells, lens, samples, n, records, stream))) return rc;
    return scar_aggregate_dev(c, seq, records, n, first_index, stream);
}

int scar_check_status(scar_ctx *c)
{
    int32_t st = 0;
    CU(cudaMemcpy(&st, c->d_status, sizeof(st), cudaMemcpyDeviceToHost));
    if (st & 1) return fail(SCAR_ERR_UNSUPPORTED, "a read is longer than max_read_len=%d", c->max_read_len);
    if (st & 2) return fail(SCAR_ERR_TABLE_FULL, "scar table full (capacity %llu)", (unsigned long long)c->capacity);
    if (st & 4) return fail(SCAR_ERR_TABLE_FULL, "an export region overflowed (raise region_capacity)");
    return SCAR_OK;
}

extern "C" int scar_verify_dev(scar_ctx *c, const scar_ragged *seq, const scar_record *records, int64_t n,
                               uint64_t first_index, int64_t *n_collisions, void *stream)
{
    if (!c || !seq || !n_collisions) return fail(SCAR_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaMemsetAsync(c->d_collisions, 0, sizeof(unsigned long long), s));
    K4VerifyParams P; memset(&P, 0, sizeof(P));
    P.seq = *seq; P.records = records; P.n_reads = n; P.first_index = first_index; P.slots = c->d_slots;
    P.slot_mask = c->capacity - 1; P.n_collisions = c->d_collisions; P.platform = c->platform; P.weak_keys = c->weak_keys ? 1 : 0;
    P.max_probe = (uint32_t)std::min<uint64_t>(c->capacity - 1, 1u << 16);
    CU(scar_launch_verify(P, c->sm_count, s));
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(&v, c->d_collisions, sizeof(v), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    *n_collisions = (int64_t)v;
    return v ? fail(SCAR_ERR_HASH_COLLISION, "%llu reads disagree with their table entry", v) : SCAR_OK;
}

// ---- results --------------------------------------------------------------------------------------
extern "C" int scar_counters_read(scar_ctx *c, int64_t *counters, int64_t *unassigned)
{
    if (!c) return fail(SCAR_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    int rc = scar_check_status(c);
    if (rc) return rc;
    if (counters) CU(cudaMemcpy(counters, c->d_counters, (size_t)c->n_samples * SCAR_N_COUNTERS * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (unassigned) CU(cudaMemcpy(unassigned, c->d_unassigned, sizeof(int64_t), cudaMemcpyDeviceToHost));
    return SCAR_OK;
}

extern "C" int scar_phase_hist_read(scar_ctx *c, int64_t *hist)
{
    if (!c || !hist) return fail(SCAR_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(hist, c->d_phase_hist, (size_t)c->n_samples * 2 * SCAR_PHASE_BINS * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return SCAR_OK;
}

extern "C" int scar_table_size(scar_ctx *c, int64_t *n)
{
    if (!c || !n) return fail(SCAR_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    int rc = scar_check_status(c);
    if (rc) return rc;
    unsigned long long v = 0;
    CU(cudaMemcpy(&v, c->d_n_entries, sizeof(v), cudaMemcpyDeviceToHost));
    *n = (int64_t)v;
    return SCAR_OK;
}

extern "C" int scar_table_export_dev(scar_ctx *c, scar_entry *entries, int64_t capacity, int64_t *n_entries, void *stream)
{
    if (!c || !n_entries || capacity < 0) return fail(SCAR_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    CU(scar_launch_export(c->d_slots, c->d_claimed, c->d_n_entries, entries, (uint64_t)capacity, c->sm_count, s));
    ++c->launches;
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(&v, c->d_n_entries, sizeof(v), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    *n_entries = (int64_t)v;
    if ((int64_t)v > capacity) return fail(SCAR_ERR_INVALID, "export buffer too small: %llu entries, capacity %lld", v, (long long)capacity);
    return SCAR_OK;
}

extern "C" int scar_table_merge_dev(scar_ctx *c, const scar_entry *entries, int64_t n, void *stream)
{
    if (!c || n < 0) return fail(SCAR_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    K4Params P = k4_params(c);
    CU(scar_launch_merge(P, entries, (uint64_t)n, c->sm_count, (cudaStream_t)stream));
    if (n) ++c->launches;
    return SCAR_OK;
}

extern "C" int scar_table_export_parts_dev(scar_ctx *c, scar_entry *regions, int32_t n_parts, int64_t region_capacity,
                                           void *stream)
{
    if (!c || !regions || n_parts < 1 || region_capacity < 2) return fail(SCAR_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    for (int g = 0; g < n_parts; ++g)        // headers (count = 0)
        CU(cudaMemsetAsync(regions + (size_t)g * region_capacity, 0, sizeof(scar_entry), s));
    CU(scar_launch_export_parts(c->d_slots, c->d_claimed, c->d_n_entries, regions, (uint32_t)n_parts, (uint64_t)region_capacity,
                                c->d_status, c->sm_count, s));
    ++c->launches;
    return SCAR_OK;
}

extern "C" int scar_table_export_p2p_dev(scar_ctx *c, void *const *peer_regions, int32_t n_parts, int32_t my_rank,
                                         int64_t region_capacity, void *stream)
{
    if (!c || !peer_regions || n_parts < 1 || my_rank < 0 || my_rank >= n_parts || region_capacity < 2)
        return fail(SCAR_ERR_INVALID, "bad argument");
    if (n_parts > SCAR_MAX_PEERS) return fail(SCAR_ERR_UNSUPPORTED, "more than %d peers", SCAR_MAX_PEERS);
    for (int o = 0; o < n_parts; ++o)
        if (!peer_regions[o]) return fail(SCAR_ERR_INVALID, "peer %d: null receive buffer", o);
    CU(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (!c->d_part_cursor) {
        CU(cudaMalloc((void **)&c->d_part_cursor, SCAR_MAX_PEERS * sizeof(unsigned long long)));
        CU(cudaMemsetAsync(c->d_part_cursor, 0, SCAR_MAX_PEERS * sizeof(unsigned long long), s));
    }
    CU(scar_launch_export_p2p(c->d_slots, c->d_claimed, c->d_n_entries, reinterpret_cast<scar_entry *const *>(peer_regions), (uint32_t)n_parts,
                              (uint32_t)my_rank, (uint64_t)region_capacity, c->d_part_cursor, c->d_status, c->sm_count, s));
    c->launches += 2;
    return SCAR_OK;
}

extern "C" int scar_table_clear_dev(scar_ctx *c, void *stream)
{
    if (!c) return fail(SCAR_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    CU(scar_launch_clear(c->d_slots, c->d_claimed, c->d_n_entries, c->sm_count, s));
    ++c->launches;
    return SCAR_OK;
}

extern "C" int scar_table_merge_parts_dev(scar_ctx *c, const scar_entry *regions, int32_t n_parts, int64_t region_capacity,
                                          void *stream)
{
    if (!c || !regions || n_parts < 1 || region_capacity < 2) return fail(SCAR_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    K4Params P = k4_params(c);
    CU(scar_launch_merge_parts(P, regions, (uint32_t)n_parts, (uint64_t)region_capacity, c->sm_count, (cudaStream_t)stream));
    ++c->launches;
    return SCAR_OK;
}

extern "C" int64_t scar_accumulator_words(const scar_ctx *c) { return c ? (int64_t)(c->acc_words - 7) : 0; }

extern "C" int scar_accumulators_export_dev(scar_ctx *c, int64_t *out, void *stream)
{
    if (!c || !out) return fail(SCAR_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaMemcpyAsync(out, c->d_unassigned, sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(out + 1, c->d_counters, (c->acc_words - 8) * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    return SCAR_OK;
}

extern "C" int scar_accumulators_import_dev(scar_ctx *c, const int64_t *in, void *stream)
{
    if (!c || !in) return fail(SCAR_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaMemcpyAsync(c->d_unassigned, in, sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(c->d_counters, in + 1, (c->acc_words - 8) * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    return SCAR_OK;
}

extern "C" int scar_table_read(scar_ctx *c, scar_entry *entries, int64_t capacity, int64_t *n_entries)
{
    if (!c || !n_entries) return fail(SCAR_ERR_INVALID, "bad argument");
    int64_t n = 0;
    int rc = scar_table_size(c, &n);
    if (rc) return rc;
    *n_entries = n;
    if (!entries) return SCAR_OK;
    if (capacity < n) return fail(SCAR_ERR_INVALID, "entries buffer too small: need %lld", (long long)n);
    if (n == 0) return SCAR_OK;
    // (device scratch kept in the context: no cudaMalloc / cudaFree per call)
    if ((rc = ensure(&c->d_gather_out, &c->gather_out_cap, (size_t)n * sizeof(scar_entry)))) return rc;
    scar_entry *d = reinterpret_cast<scar_entry *>(c->d_gather_out);
    int64_t got = 0;
    rc = scar_table_export_dev(c, d, n, &got, nullptr);
    std::vector<scar_entry> tmp;
    if (!rc) {
        tmp.resize((size_t)n);
        cudaError_t e = cudaMemcpy(tmp.data(), d, (size_t)n * sizeof(scar_entry), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(SCAR_ERR_CUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
    }
    if (rc) return rc;
    // results_freq_dict order: per sample, first-seen (INDEL_Processing.py:193-196).  Counting sort by sample, then
    // every sample's entries by first index, the samples spread over a few threads.
    std::vector<int64_t> start((size_t)c->n_samples + 2, 0);
    for (int64_t k = 0; k < n; ++k) {
        if ((int64_t)tmp[(size_t)k].sample > c->n_samples) return fail(SCAR_ERR_CUDA, "table entry with sample %u", (unsigned)tmp[(size_t)k].sample);
        ++start[(size_t)tmp[(size_t)k].sample + 1];
    }
    for (size_t s = 1; s < start.size(); ++s) start[s] += start[s - 1];
    {
        std::vector<int64_t> cursor(start.begin(), start.end() - 1);
        for (int64_t k = 0; k < n; ++k) entries[cursor[(size_t)tmp[(size_t)k].sample]++] = tmp[(size_t)k];
    }
    const int n_buckets = (int)start.size() - 1;
    const int threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(8, (int64_t)std::thread::hardware_concurrency()), n / 65536));
    std::atomic<int> next(0);
    auto work = [&]() {
        for (int b = next++; b < n_buckets; b = next++)
            std::sort(entries + start[(size_t)b], entries + start[(size_t)b + 1],
                      [](const scar_entry &x, const scar_entry &y) { return x.first_idx < y.first_idx; });
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    return SCAR_OK;
}

// ---- host-buffer pipeline -------------------------------------------------------------------------
struct Span { int64_t byte0, byte1; };   // host byte range of items [a, b)
static Span span_of(const scar_ragged &R, int64_t a, int64_t b)
{
    if (b <= a) return {0, 0};
    if (R.offsets && R.lengths) return {R.offsets[a], R.offsets[b - 1] + R.lengths[b - 1]};   // views, monotone offsets
    if (R.offsets) return {R.offsets[a], R.offsets[b]};
    return {a * R.stride, b * R.stride};
}

// copy the index arrays of items [a,b) to the device and point the descriptor at device bytes that hold
// host bytes [byte0, ...) at dbytes
static int stage_index(const scar_ragged &H, int64_t a, int64_t b, const uint8_t *dbytes, int64_t byte0,
                       int64_t *doff, int32_t *dlen, cudaStream_t s, scar_ragged *D, int64_t *h2d)
{
    memset(D, 0, sizeof(*D));
    D->stride = H.stride;
    if (H.offsets) {
        const size_t n_off = (size_t)(b - a) + (H.lengths ? 0 : 1);
        CU(cudaMemcpyAsync(doff, H.offsets + a, n_off * sizeof(int64_t), cudaMemcpyHostToDevice, s));
        *h2d += (int64_t)(n_off * sizeof(int64_t));
        D->offsets = doff;
        D->bytes = dbytes - byte0;           // original offsets stay valid
    } else {
        D->bytes = dbytes;
    }
    if (H.lengths) {
        CU(cudaMemcpyAsync(dlen, H.lengths + a, (size_t)(b - a) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        *h2d += (int64_t)((b - a) * sizeof(int32_t));
        D->lengths = dlen;
    }
    return SCAR_OK;
}

// copy items [a,b) of a host ragged list to the device; returns the device-side descriptor
static int stage_ragged(const scar_ragged &H, int64_t a, int64_t b, uint8_t **dbytes, size_t *dcap,
                        int64_t **doff, int32_t **dlen, int64_t, cudaStream_t s, scar_ragged *D,
                        int64_t *h2d)
{
    const Span sp = span_of(H, a, b);
    const size_t nbytes = (size_t)(sp.byte1 - sp.byte0);
    int rc = ensure(dbytes, dcap, nbytes + 16);
    if (rc) return rc;
    if (nbytes) CU(cudaMemcpyAsync(*dbytes, H.bytes + sp.byte0, nbytes, cudaMemcpyHostToDevice, s));
    *h2d += (int64_t)nbytes;
    return stage_index(H, a, b, *dbytes, H.offsets ? sp.byte0 : 0, *doff, *dlen, s, D, h2d);
}

int scar_ensure_hostbuf(scar_ctx *c, HostBuf &b, int64_t reads)
{
    if (!b.stream) { CU(cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking)); CU(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming)); }
    if (reads <= b.reads_cap) return SCAR_OK;
    cudaFree(b.seq_off); cudaFree(b.f0_off); cudaFree(b.f1_off); cudaFree(b.seq_len); cudaFree(b.f0_len); cudaFree(b.f1_len);
    cudaFree(b.cells); cudaFree(b.lens); cudaFree(b.samples); cudaFree(b.records);
    b.seq_off = b.f0_off = b.f1_off = nullptr; b.seq_len = b.f0_len = b.f1_len = nullptr;
    b.cells = nullptr; b.lens = nullptr; b.samples = nullptr; b.records = nullptr;
    b.reads_cap = 0;
    const size_t n = (size_t)reads;
    CU(cudaMalloc((void **)&b.seq_off, (n + 1) * sizeof(int64_t)));
    CU(cudaMalloc((void **)&b.f0_off, (n + 1) * sizeof(int64_t)));
    CU(cudaMalloc((void **)&b.f1_off, (n + 1) * sizeof(int64_t)));
    CU(cudaMalloc((void **)&b.seq_len, n * sizeof(int32_t)));
    CU(cudaMalloc((void **)&b.f0_len, n * sizeof(int32_t)));
    CU(cudaMalloc((void **)&b.f1_len, n * sizeof(int32_t)));
    if (!scar_fused_path(c)) {                                  // the packed batch exists only on the K1 -> K3 -> K4 path
        CU(cudaMalloc((void **)&b.cells, n * c->W * sizeof(uint64_t)));
        CU(cudaMalloc((void **)&b.lens, n * sizeof(uint32_t)));
    }
    CU(cudaMalloc((void **)&b.samples, n * sizeof(uint16_t)));
    CU(cudaMalloc((void **)&b.records, n * sizeof(scar_record)));
    b.reads_cap = reads;
    return SCAR_OK;
}

extern "C" int scar_process_host(scar_ctx *c, const scar_ragged *seq, const scar_ragged *f0, const scar_ragged *f1,
                                 int64_t n, uint64_t first_index, scar_record *records)
{
    if (!c || !seq || n < 0) return fail(SCAR_ERR_INVALID, "bad argument");
    if (c->platform == SCAR_PLATFORM_ILLUMINA && (!f0 || !f1)) return fail(SCAR_ERR_INVALID, "Illumina needs the two header index fields");
    CU(cudaSetDevice(c->device));
    // verify mode: the whole batch is ONE resident chunk, so that every key's first read is on the device when
    // scar_verify_dev compares each scarred read's full key with it (exact-or-fail identity instead of 128 hash bits)
    const int64_t chunk = c->verify ? std::max<int64_t>(n, 1) : std::min<int64_t>(c->chunk_reads, std::max<int64_t>(n, 1));
    int k = 0;
    int64_t h2d = 0;
    for (int64_t a = 0; a < n; a += chunk, ++k) {
        const int64_t b = std::min(n, a + chunk);
        HostBuf &B = c->buf[k % NBUF];
        int rc = scar_ensure_hostbuf(c, B, chunk);
        if (rc) return rc;
        if (k >= NBUF) CU(cudaEventSynchronize(B.done));     // buffer reuse: its previous chunk has drained
        scar_ragged dseq, df0, df1;
        const bool illumina = c->platform == SCAR_PLATFORM_ILLUMINA;
        const bool shared = illumina && seq->offsets && f0->offsets && f1->offsets && f0->bytes == seq->bytes &&
                            f1->bytes == seq->bytes;
        if (shared) {
            // the three lists are views into ONE text buffer (FASTQ text): copy its span once
            const Span s0 = span_of(*seq, a, b), s1 = span_of(*f0, a, b), s2 = span_of(*f1, a, b);
            const int64_t lo = std::min(s0.byte0, std::min(s1.byte0, s2.byte0));
            const int64_t hi = std::max(s0.byte1, std::max(s1.byte1, s2.byte1));
            if ((rc = ensure(&B.seq, &B.seq_cap, (size_t)(hi - lo) + 16))) return rc;
            if (hi > lo) CU(cudaMemcpyAsync(B.seq, seq->bytes + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, B.stream));
            h2d += hi - lo;
            if ((rc = stage_index(*seq, a, b, B.seq, lo, B.seq_off, B.seq_len, B.stream, &dseq, &h2d))) return rc;
            if ((rc = stage_index(*f0, a, b, B.seq, lo, B.f0_off, B.f0_len, B.stream, &df0, &h2d))) return rc;
            if ((rc = stage_index(*f1, a, b, B.seq, lo, B.f1_off, B.f1_len, B.stream, &df1, &h2d))) return rc;
        } else {
            if ((rc = stage_ragged(*seq, a, b, &B.seq, &B.seq_cap, &B.seq_off, &B.seq_len, 0, B.stream, &dseq, &h2d))) return rc;
            if (illumina) {
                if ((rc = stage_ragged(*f0, a, b, &B.f0, &B.f0_cap, &B.f0_off, &B.f0_len, 0, B.stream, &df0, &h2d))) return rc;
                if ((rc = stage_ragged(*f1, a, b, &B.f1, &B.f1_cap, &B.f1_off, &B.f1_len, 0, B.stream, &df1, &h2d))) return rc;
            }
        }
        if ((rc = scar_run_dev(c, &dseq, c->platform == SCAR_PLATFORM_ILLUMINA ? &df0 : nullptr,
                               c->platform == SCAR_PLATFORM_ILLUMINA ? &df1 : nullptr, b - a, first_index + (uint64_t)a,
                               B.cells, B.lens, B.samples, B.records, B.stream))) return rc;
        if (records)
            CU(cudaMemcpyAsync(records + a, B.records, (size_t)(b - a) * sizeof(scar_record), cudaMemcpyDeviceToHost, B.stream));
        if (c->verify) {
            int64_t ncol = 0;
            if ((rc = scar_verify_dev(c, &dseq, B.records, b - a, first_index + (uint64_t)a, &ncol, B.stream))) return rc;
        }
        CU(cudaEventRecord(B.done, B.stream));
    }
    for (int i = 0; i < NBUF; ++i) if (c->buf[i].stream) CU(cudaStreamSynchronize(c->buf[i].stream));
    return scar_check_status(c);
}

// ---- FASTQ text in host memory -> records: the text is copied to the GPU once and indexed THERE ----------
extern "C" int scar_process_fastq_host(scar_ctx *c, const uint8_t *text, int64_t n_bytes, uint64_t first_index,
                                       scar_record *records, int64_t records_capacity, int64_t *n_records_out)
{
    if (!c || !text || n_bytes < 0 || !n_records_out) return fail(SCAR_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    *n_records_out = 0;
    c->fq_records_n = 0;
    if (n_bytes == 0) return SCAR_OK;
    HostBuf &B = c->buf[0];
    int rc = scar_ensure_hostbuf(c, B, 1);
    if (rc) return rc;
    if ((rc = scar_ensure_hostbuf(c, c->buf[1], 1))) return rc;
    cudaStream_t s = B.stream, sc = c->buf[1].stream;
    if ((rc = ensure(&c->d_text, &c->text_cap, (size_t)n_bytes + 64))) return rc;
    // The text is copied in pieces of plain bytes on one stream; a second stream indexes every piece as soon as
    // it has landed (newline count, scan, positions - all global offsets into the one device buffer) and runs
    // the records that have become complete through K2/K1/K3/K4, so the kernels hide under the copy of the
    // next piece.  A record may straddle pieces: it is simply picked up with the piece that completes it.
    int64_t piece = 256ll << 20;
    if (const char *e = getenv("SCAR_FASTQ_PIECE_BYTES")) { const long long v = atoll(e); if (v > 0) piece = v; }
    piece = std::max<int64_t>(4096, (piece + 4095) / 4096 * 4096);
    const int64_t n_pieces = (n_bytes + piece - 1) / piece;
    while ((int64_t)c->fq_events.size() < n_pieces) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->fq_events.push_back(e);
    }
    for (int64_t k = 0; k < n_pieces; ++k) {
        const int64_t b0 = k * piece, len = std::min(piece, n_bytes - b0);
        CU(cudaMemcpyAsync(c->d_text + b0, text + b0, (size_t)len, cudaMemcpyHostToDevice, sc));
        if (k == n_pieces - 1) CU(cudaMemsetAsync(c->d_text + n_bytes, 0, 64, sc));
        CU(cudaEventRecord(c->fq_events[k], sc));
    }
    if ((rc = ensure(&c->d_tile_counts, &c->tile_counts_cap, (size_t)((piece + 4095) / 4096)))) return rc;
    if ((rc = ensure(&c->d_tile_offsets, &c->tile_offsets_cap, (size_t)scar_fastq_offsets_elems(piece)))) return rc;
    const int want = c->platform == SCAR_PLATFORM_ILLUMINA;
    int64_t nl_total = 0, rec_done = 0;
    // an early return must not leave copies out of the caller's buffer in flight
    auto drain = [&](int code) { cudaStreamSynchronize(sc); cudaStreamSynchronize(s); return code; };
    for (int64_t k = 0; k < n_pieces; ++k) {
        const int64_t b0 = k * piece, len = std::min(piece, n_bytes - b0);
        CU(cudaStreamWaitEvent(s, c->fq_events[k], 0));
        CU(scar_fastq_index_launch(c->d_text + b0, len, c->d_tile_counts, c->d_tile_offsets, c->d_export_n, s));
        c->launches += 3;
        unsigned long long n_nl = 0;
        CU(cudaMemcpyAsync(&n_nl, c->d_export_n, sizeof(n_nl), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        // newline positions of the whole text, grown from the density seen so far
        size_t want_nl = (size_t)(nl_total + (int64_t)n_nl) + 1;
        if (want_nl > c->nl_cap && k + 1 < n_pieces)
            want_nl = (size_t)((double)want_nl * (double)n_bytes / (double)(b0 + len) * 1.02) + 1024;
        if ((rc = ensure_keep(&c->d_nl, &c->nl_cap, want_nl, (size_t)nl_total))) return drain(rc);
        CU(scar_fastq_positions_launch(c->d_text + b0, len, c->d_tile_offsets, c->d_nl + nl_total, (int64_t)n_nl, b0, s));
        ++c->launches;
        nl_total += (int64_t)n_nl;
        // whole records only; the last record may lack its final newline
        int64_t rec_avail = nl_total / 4;
        if (k == n_pieces - 1 && nl_total % 4 == 3 && text[n_bytes - 1] != '\n') ++rec_avail;
        const int64_t m = rec_avail - rec_done;
        if (m <= 0) continue;
        if (records && records_capacity < rec_avail)
            return drain(fail(SCAR_ERR_INVALID, "records buffer too small: more than %lld records", (long long)records_capacity));
        if ((rc = scar_ensure_hostbuf(c, B, m))) return drain(rc);
        size_t want_rec = (size_t)rec_avail;
        if (want_rec > c->fq_records_cap && k + 1 < n_pieces)
            want_rec = (size_t)((double)want_rec * (double)n_bytes / (double)(b0 + len) * 1.02) + 1024;
        if ((rc = ensure_keep(&c->d_fq_records, &c->fq_records_cap, want_rec, (size_t)rec_done))) return drain(rc);
        CU(scar_fastq_records_launch(c->d_text, n_bytes, c->d_nl, nl_total, rec_done, m, want, B.seq_off, B.seq_len, B.f0_off,
                                     B.f0_len, B.f1_off, B.f1_len, c->d_status, c->sm_count, s));
        ++c->launches;
        scar_ragged dseq{c->d_text, B.seq_off, B.seq_len, 0}, df0{c->d_text, B.f0_off, B.f0_len, 0}, df1{c->d_text, B.f1_off, B.f1_len, 0};
        if ((rc = scar_run_dev(c, &dseq, want ? &df0 : nullptr, want ? &df1 : nullptr, m, first_index + (uint64_t)rec_done,
                               B.cells, B.lens, B.samples, c->d_fq_records + rec_done, s))) return drain(rc);
        if (records)
            CU(cudaMemcpyAsync(records + rec_done, c->d_fq_records + rec_done, (size_t)m * sizeof(scar_record),
                               cudaMemcpyDeviceToHost, s));
        rec_done = rec_avail;
    }
    CU(cudaStreamSynchronize(s));
    CU(cudaStreamSynchronize(sc));
    *n_records_out = rec_done;
    c->fq_records_n = rec_done;
    int32_t st = 0;
    CU(cudaMemcpy(&st, c->d_status, sizeof(st), cudaMemcpyDeviceToHost));
    if (st & 24) {
        CU(cudaMemset(c->d_status, 0, sizeof(int32_t)));
        if (st & 8) return fail(SCAR_ERR_INVALID, "FASTQ: sequence and quality scores of different lengths");
        return fail(SCAR_ERR_INVALID, "FASTQ: a header has no '+' between the index reads (the reference raises IndexError)");
    }
    return scar_check_status(c);
}

extern "C" int scar_fastq_records_fetch(scar_ctx *c, scar_record *records, int64_t n_records)
{
    if (!c || !records || n_records < 0) return fail(SCAR_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    if (n_records > c->fq_records_n) return fail(SCAR_ERR_INVALID, "no such records on the device");
    if (n_records) CU(cudaMemcpy(records, c->d_fq_records, (size_t)n_records * sizeof(scar_record), cudaMemcpyDeviceToHost));
    return SCAR_OK;
}

// ---- Sequence_Magic.match_maker batch -------------------------------------------------------------
extern "C" int scar_match_maker_batch(int32_t device, const scar_ragged *query, const scar_ragged *unknown, int64_t n, int32_t *out)
{
    if (!query || !unknown || !out || n < 0) return fail(SCAR_ERR_INVALID, "bad argument");
    if (n == 0) return SCAR_OK;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(SCAR_ERR_CUDA, "CUDA device %d not present", device);
    CU(cudaSetDevice(device));
    for (int64_t i = 0; i < n; ++i) {
        const int64_t m = query->lengths ? query->lengths[i] : (query->offsets ? query->offsets[i + 1] - query->offsets[i] : query->stride);
        if (m > 64) return fail(SCAR_ERR_UNSUPPORTED, "query %lld longer than 64", (long long)i);
    }
    uint8_t *dq = nullptr, *du = nullptr; int64_t *dqo = nullptr, *duo = nullptr; int32_t *dql = nullptr, *dul = nullptr, *dout = nullptr;
    size_t cq = 0, cu_ = 0;
    int rc = SCAR_OK;
    int64_t h2d = 0;
    scar_ragged Dq, Du;
    auto cleanup = [&]() { cudaFree(dq); cudaFree(du); cudaFree(dqo); cudaFree(duo); cudaFree(dql); cudaFree(dul); cudaFree(dout); };
    if (cudaMalloc((void **)&dqo, (n + 1) * sizeof(int64_t)) != cudaSuccess || cudaMalloc((void **)&duo, (n + 1) * sizeof(int64_t)) != cudaSuccess ||
        cudaMalloc((void **)&dql, n * sizeof(int32_t)) != cudaSuccess || cudaMalloc((void **)&dul, n * sizeof(int32_t)) != cudaSuccess ||
        cudaMalloc((void **)&dout, n * sizeof(int32_t)) != cudaSuccess) { cleanup(); return fail(SCAR_ERR_NOMEM, "cudaMalloc failed"); }
    if (!(rc = stage_ragged(*query, 0, n, &dq, &cq, &dqo, &dql, 0, nullptr, &Dq, &h2d)) &&
        !(rc = stage_ragged(*unknown, 0, n, &du, &cu_, &duo, &dul, 0, nullptr, &Du, &h2d))) {
        cudaError_t e = scar_launch_match_maker(Dq, Du, n, dout, nullptr);
        if (e == cudaSuccess) e = cudaMemcpy(out, dout, n * sizeof(int32_t), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(SCAR_ERR_CUDA, "match_maker kernel failed: %s", cudaGetErrorString(e));
    }
    cleanup();
    return rc;
}

// ---- homeology / templated-insertion searches over a batch of scar patterns -------------------------
namespace {
struct StagedList {                 // one host ragged list copied to the device
    uint8_t *bytes = nullptr; size_t cap = 0; int64_t *off = nullptr; int32_t *len = nullptr; scar_ragged dev{};
    ~StagedList() { cudaFree(bytes); cudaFree(off); cudaFree(len); }
    int stage(const scar_ragged &H, int64_t n)
    {
        int64_t h2d = 0;
        if (cudaMalloc((void **)&off, (size_t)(n + 1) * sizeof(int64_t)) != cudaSuccess ||
            cudaMalloc((void **)&len, (size_t)std::max<int64_t>(n, 1) * sizeof(int32_t)) != cudaSuccess)
            return fail(SCAR_ERR_NOMEM, "cudaMalloc failed");
        return stage_ragged(H, 0, n, &bytes, &cap, &off, &len, 0, nullptr, &dev, &h2d);
    }
};
struct DevInts {
    int32_t *p = nullptr;
    ~DevInts() { cudaFree(p); }
    int upload(const int32_t *h, int64_t n)
    {
        if (cudaMalloc((void **)&p, (size_t)std::max<int64_t>(n, 1) * sizeof(int32_t)) != cudaSuccess)
            return fail(SCAR_ERR_NOMEM, "cudaMalloc failed");
        CU(cudaMemcpy(p, h, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
        return SCAR_OK;
    }
};
}   // namespace

namespace {
// One grow-only block of device memory per device for the batch entry points that have no context: a call carves
// its buffers out of it.  cudaMalloc / cudaFree per buffer cost more than the kernels here, and cudaFree stalls
// unpredictably (hundreds of ms were measured inside a pipeline run) - nothing is freed on the way out.
struct DeviceArena {
    uint8_t *base = nullptr; size_t cap = 0, used = 0;
    int reserve(size_t total)
    {
        used = 0;
        if (total <= cap) return SCAR_OK;
        if (base) cudaFree(base);
        base = nullptr; cap = 0;
        const size_t want = total + total / 8 + (1 << 20);
        if (cudaMalloc((void **)&base, want) != cudaSuccess) { cudaGetLastError(); base = nullptr; return fail(SCAR_ERR_NOMEM, "no device memory for %zu bytes", want); }
        cap = want;
        return SCAR_OK;
    }
    template <typename T> T *take(size_t count) { T *p = reinterpret_cast<T *>(base + used); used += (count * sizeof(T) + 255) & ~(size_t)255; return p; }
};
std::mutex g_arena_mutex;
DeviceArena g_arena[64];

size_t ragged_device_bytes(const scar_ragged &H, int64_t n)
{
    const Span sp = span_of(H, 0, n);
    return (size_t)(sp.byte1 - sp.byte0) + 16 + 256 + ((size_t)(n + 1) * 8 + 256) + ((size_t)std::max<int64_t>(n, 1) * 4 + 256);
}

int stage_from_arena(DeviceArena &A, const scar_ragged &H, int64_t n, scar_ragged *D)
{
    const Span sp = span_of(H, 0, n);
    size_t cap = (size_t)(sp.byte1 - sp.byte0) + 16;
    uint8_t *bytes = A.take<uint8_t>(cap);
    int64_t *off = A.take<int64_t>((size_t)n + 1);
    int32_t *len = A.take<int32_t>((size_t)std::max<int64_t>(n, 1));
    int64_t h2d = 0;
    return stage_ragged(H, 0, n, &bytes, &cap, &off, &len, 0, nullptr, D, &h2d);      // (cap suffices: nothing is allocated)
}
}   // namespace

extern "C" int scar_homeology_batch(int32_t device, const scar_ragged *consensus, const scar_ragged *insertion,
                                    const int32_t *clj, const int32_t *crj, const int32_t *tlj, const int32_t *trj,
                                    const int32_t *target_idx, const scar_ragged *targets, const scar_ragged *regions,
                                    int32_t n_targets, int64_t n, int32_t *out)
{
    if (!consensus || !insertion || !clj || !crj || !tlj || !trj || !target_idx || !targets || !regions || !out || n < 0 ||
        n_targets < 0)
        return fail(SCAR_ERR_INVALID, "bad argument");
    if (n == 0) return SCAR_OK;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev || device >= 64) return fail(SCAR_ERR_CUDA, "CUDA device %d not present", device);
    CU(cudaSetDevice(device));
    const bool trace = getenv("SCAR_TRACE") != nullptr;
    auto now = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; };
    const double t0 = now();
    std::lock_guard<std::mutex> lock(g_arena_mutex);
    DeviceArena &A = g_arena[device];
    const size_t ints = ((size_t)n * 4 + 256);
    int rc = A.reserve(ragged_device_bytes(*consensus, n) + ragged_device_bytes(*insertion, n) + ragged_device_bytes(*targets, n_targets) +
                       ragged_device_bytes(*regions, n_targets) + 5 * ints + (size_t)n * 64 + 256);
    if (rc) return rc;
    scar_ragged dc, di, dt, dr;
    if ((rc = stage_from_arena(A, *consensus, n, &dc)) || (rc = stage_from_arena(A, *insertion, n, &di)) ||
        (rc = stage_from_arena(A, *targets, n_targets, &dt)) || (rc = stage_from_arena(A, *regions, n_targets, &dr)))
        return rc;
    int32_t *dint[5];
    const int32_t *hint[5] = {clj, crj, tlj, trj, target_idx};
    for (int k = 0; k < 5; ++k) {
        dint[k] = A.take<int32_t>((size_t)n);
        CU(cudaMemcpy(dint[k], hint[k], (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    int32_t *dout = A.take<int32_t>((size_t)n * 16);
    if (trace) cudaDeviceSynchronize();
    const double t1 = now();
    cudaError_t err = scar_launch_homeology(dc, di, dt, dr, dint[0], dint[1], dint[2], dint[3], dint[4], n_targets, n, dout, nullptr);
    if (err == cudaSuccess && trace) err = cudaDeviceSynchronize();
    const double t2 = now();
    if (err == cudaSuccess) err = cudaMemcpy(out, dout, (size_t)n * 16 * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (trace)
        fprintf(stderr, "[scar homeology] %lld patterns: staged %.1f ms, kernel %.1f ms, results %.1f ms (device arena %zu MB)\n",
                (long long)n, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (now() - t2), A.cap >> 20);
    if (err != cudaSuccess) return fail(SCAR_ERR_CUDA, "homeology kernel failed: %s", cudaGetErrorString(err));
    return SCAR_OK;
}

// ---- AlignmentProcessing.alignment_processing over a batch of gapped alignments -------------------------
extern "C" int scar_alignment_batch(int32_t device, const scar_ragged *genomic, const scar_ragged *consensus,
                                    const int32_t *cutposition, const int32_t *gap_con_len, const int32_t *gap_genome_len,
                                    int64_t n, int32_t max_spans, int32_t *counts, int32_t *spans)
{
    if (!genomic || !consensus || !cutposition || !gap_con_len || !gap_genome_len || !counts || !spans || n < 0 || max_spans < 1)
        return fail(SCAR_ERR_INVALID, "bad argument");
    if (n == 0) return SCAR_OK;
    // the reference indexes both strings up to gap_con_len (an IndexError there): the batch takes well-formed
    // alignments only - two strings of one length, gap_con_len equal to it
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t *g, *c; int ng, nc;
        get_item(*genomic, i, g, ng);
        get_item(*consensus, i, c, nc);
        if (ng != nc || gap_con_len[i] != nc)
            return fail(SCAR_ERR_UNSUPPORTED, "alignment %lld: the two gapped strings and gap_con_len must have one length", (long long)i);
    }
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(SCAR_ERR_CUDA, "CUDA device %d not present", device);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    StagedList dg, dc;
    DevInts a, b, c;
    int rc;
    if ((rc = dg.stage(*genomic, n)) || (rc = dc.stage(*consensus, n)) || (rc = a.upload(cutposition, n)) ||
        (rc = b.upload(gap_con_len, n)) || (rc = c.upload(gap_genome_len, n)))
        return rc;
    int32_t *dcounts = nullptr, *dspans = nullptr;
    const size_t span_ints = (size_t)n * 3 * max_spans * 2;
    if (cudaMalloc((void **)&dcounts, (size_t)n * 8 * sizeof(int32_t)) != cudaSuccess ||
        cudaMalloc((void **)&dspans, span_ints * sizeof(int32_t)) != cudaSuccess) {
        cudaFree(dcounts);
        return fail(SCAR_ERR_NOMEM, "cudaMalloc failed");
    }
    cudaError_t err = cudaMemset(dspans, 0, span_ints * sizeof(int32_t));
    if (err == cudaSuccess)
        err = scar_launch_alignment(dg.dev, dc.dev, a.p, b.p, c.p, n, max_spans, dcounts, dspans, prop.multiProcessorCount, nullptr);
    if (err == cudaSuccess) err = cudaMemcpy(counts, dcounts, (size_t)n * 8 * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (err == cudaSuccess) err = cudaMemcpy(spans, dspans, span_ints * sizeof(int32_t), cudaMemcpyDeviceToHost);
    cudaFree(dcounts); cudaFree(dspans);
    if (err != cudaSuccess) return fail(SCAR_ERR_CUDA, "alignment kernel failed: %s", cudaGetErrorString(err));
    return SCAR_OK;
}
