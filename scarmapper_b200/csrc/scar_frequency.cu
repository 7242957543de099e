// scar_frequency.cu — the per-pattern body of ScarSearch.frequency_output (reference scarmapper/INDEL_Processing.py:
// 233-478) for one sample, host side: scar-type classification, junction-type totals, SNV tallies, PatternThreshold
// filter, ordering and the text rows of the *_ScarMapper_Frequency.txt file.  The reference does this in a Python
// loop over every unique scar pattern (string slicing, reverse complements, a natsorted dict, str() of every field);
// with ~1 M patterns per run that loop is what remains of the wall time once the reads are classified on the GPU
// (SURVEY.md 8f, F1).  Inputs are what the GPU pass leaves behind per sample: pattern counts in first-seen order,
// the record and stored sequence of each pattern's first read, and the homeology / templated-insertion results of
// scar_homeology_batch.  Formatting and file output are host work; the Frequency column is Python's
// str(count / total_scars), reproduced here (shortest round-trip digits, Python's fixed / exponent rule).
#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "scar_internal.cuh"
#include "scar_homeology_core.h"

int scar_fail(int code, const char *fmt, ...);

namespace {

// repr(float) of CPython (float_repr_style 'short'): shortest digits that round-trip, exponent form when the
// decimal point would lie more than 16 digits right or 4 digits left of the first digit
std::string py_float_repr(double x)
{
    if (x == 0.0) return std::signbit(x) ? "-0.0" : "0.0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);
    std::string s(buf, r.ptr);                     // [-]d[.ddd]e[+-]XX
    std::string sign;
    if (s[0] == '-') { sign = "-"; s.erase(0, 1); }
    const size_t e = s.find('e');
    std::string mant = s.substr(0, e);
    const int exp10 = atoi(s.c_str() + e + 1);
    std::string digits;
    for (char c : mant) if (c != '.') digits.push_back(c);
    const int decpt = exp10 + 1;                   // value = 0.DIGITS x 10^decpt
    const int nd = (int)digits.size();
    std::string out = sign;
    if (decpt > 16 || decpt <= -4) {               // exponent form: d[.ddd]e[+-]XX
        out.push_back(digits[0]);
        if (nd > 1) { out.push_back('.'); out.append(digits, 1, std::string::npos); }
        char eb[16];
        snprintf(eb, sizeof eb, "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
        out += eb;
    } else if (decpt <= 0) {
        out += "0.";
        out.append((size_t)(-decpt), '0');
        out += digits;
    } else if (decpt >= nd) {
        out += digits;
        out.append((size_t)(decpt - nd), '0');
        out += ".0";
    } else {
        out.append(digits, 0, (size_t)decpt);
        out.push_back('.');
        out.append(digits, (size_t)decpt, std::string::npos);
    }
    return out;
}

std::string rcomp(const std::string &s)
{
    std::string r(s.size(), 'N');
    for (size_t i = 0; i < s.size(); ++i) r[i] = (char)hom_comp((uint8_t)s[s.size() - 1 - i]);
    return r;
}

// s[a:b] with Python's clipping (a, b >= 0 here)
std::string pyslice(const std::string &s, long a, long b)
{
    const long n = (long)s.size();
    if (a < 0) a = 0;
    if (b > n) b = n;
    return b > a ? s.substr((size_t)a, (size_t)(b - a)) : std::string();
}

const char *const TYPE_NAMES[7] = {"HR", "SNV", "TMEJ", "NHEJ", "Non-MH Deletion", "Insertion", "Other"};
enum { T_HR = 0, T_SNV, T_TMEJ, T_NHEJ, T_NONMH, T_INS, T_OTHER };

struct Row {                 // what the sorted pass needs of a pattern
    uint32_t count; int32_t type, lft_del, rt_del, del_size, mh_size, ins_size;
    double key_frequency;
};

}   // namespace

static thread_local std::string g_rows;       // the rows of the calling thread's last scar_frequency_rows_host

extern "C" int scar_frequency_rows_fetch(uint8_t *rows, int64_t capacity)
{
    if (!rows || capacity < (int64_t)g_rows.size()) return scar_fail(SCAR_ERR_INVALID, "frequency rows: %lld bytes needed", (long long)g_rows.size());
    memcpy(rows, g_rows.data(), g_rows.size());
    g_rows.clear(); g_rows.shrink_to_fit();
    return SCAR_OK;
}

extern "C" int scar_frequency_rows_host(const scar_frequency_in *in, int64_t *rows_bytes,
                                        int64_t *junction_type_data, double *label_sums, int32_t *label_order, int64_t *snv,
                                        double *plot, int64_t plot_capacity, int64_t *n_plot, int64_t *totals)
{
    if (!in || !rows_bytes || !junction_type_data || !label_sums || !label_order || !snv || !n_plot || !totals || in->n < 0 ||
        (in->n && (!in->counts || !in->first || !in->reads || !in->read_off || !in->pattern)) || !in->region)
        return scar_fail(SCAR_ERR_INVALID, "bad argument");
    const int64_t n = in->n;
    const std::string region((const char *)in->region, (size_t)in->region_len);
    const bool rc = in->reverse_comp != 0;
    const int cut = in->cutsite, L = in->region_len;
    const std::string oriented = rc ? rcomp(region) : region;                 // target_sequence of the rows (:300,319)
    // left_target_windows[0] / right_target_windows[0] (window_mapping, :57-80), swapped and reverse-complemented for
    // 3'-strand guides (:315-316)
    const std::string lw0 = pyslice(region, cut - 10, cut), rw0 = pyslice(region, cut, cut + 10);
    const std::string lft_kmer = rc ? rcomp(rw0) : lw0, rt_kmer = rc ? rcomp(lw0) : rw0;
    const int kmer_size = (int)lw0.size();                                    // len(self.left_target_windows[0]) (:269)
    for (int i = 0; i < 6; ++i) junction_type_data[i] = 0;
    for (int i = 0; i < 7; ++i) { label_sums[i] = 0.0; label_order[i] = -1; }
    for (int i = 0; i < 4 * 10 * 4; ++i) snv[i] = 0;
    const double reads_passing = (double)(in->passing - in->no_cut);
    int64_t total_scars = 0;
    for (int64_t k = 0; k < n; ++k) total_scars += in->counts[k];
    totals[0] = n; totals[1] = total_scars;

    // ---- pass 1 (first-seen order, :277-394): classification, junction-type totals, SNV tallies
    std::vector<Row> R((size_t)n);
    auto nt_index = [](char c) { return c == 'G' ? 0 : c == 'A' ? 1 : c == 'T' ? 2 : c == 'C' ? 3 : -1; };   // nt_list order (:491)
    for (int64_t k = 0; k < n; ++k) {
        const scar_record &rec = in->first[k];
        Row &r = R[(size_t)k];
        r.count = in->counts[k];
        r.key_frequency = reads_passing > 0 ? (double)r.count / reads_passing : 0.0;
        const int tlj = rec.tlj, trj = rec.trj, clj = rec.clj, crj = rec.crj;
        int lft = cut > tlj ? cut - tlj : 0;                                  // len(target[tlj:cutsite])
        int rt = std::min(trj, L) > cut ? std::min(trj, L) - cut : 0;         // len(target[cutsite:trj])
        const int rlen = (int)(in->read_off[k + 1] - in->read_off[k]);
        r.ins_size = (rec.flags & SCAR_FL_INSERTION) ? std::max(0, std::min(crj, rlen) - std::min(clj, rlen)) : 0;
        r.mh_size = (rec.flags & SCAR_FL_MICROHOM) ? std::max(0, std::min(clj, rlen) - std::min(crj, rlen)) : 0;
        r.del_size = lft + rt + r.mh_size;
        if (rc) std::swap(lft, rt);
        r.lft_del = lft; r.rt_del = rt;
        if (rec.flags & SCAR_FL_HR) r.type = T_HR;
        else if (r.del_size > 1 && r.del_size == r.ins_size && lft <= kmer_size && rt <= kmer_size) {
            // SNV tallies (:333-350); the insertion in the orientation of the rows
            const char *rd = (const char *)in->reads + in->read_off[k];
            std::string ins(rd + std::min(clj, rlen), (size_t)r.ins_size);
            if (rc) ins = rcomp(ins);
            if (lft > 1) {
                int position = 0;
                for (int i = kmer_size - lft; i < kmer_size; ++i, ++position) {
                    if (ins[(size_t)position] != lft_kmer[(size_t)i]) {
                        const int q = nt_index(ins[(size_t)position]);
                        if (q >= 0 && i < 10) { snv[(0 * 10 + i) * 4 + q] += 1; snv[(1 * 10 + i) * 4 + q] += r.count; }
                    }
                }
            }
            if (rt > 1) {
                const int position = (int)ins.size() - rt;                    // (:345-349: never advanced inside the loop)
                for (int i = 0; i < rt; ++i) {
                    if (ins[(size_t)position] != rt_kmer[(size_t)i]) {
                        const int q = nt_index(ins[(size_t)position]);
                        if (q >= 0 && i < 10) { snv[(2 * 10 + i) * 4 + q] += 1; snv[(3 * 10 + i) * 4 + q] += r.count; }
                    }
                }
            }
            junction_type_data[5] += r.count;
            r.type = T_SNV;
        } else if (r.del_size >= 4 && r.mh_size >= 2) { junction_type_data[0] += r.count; r.type = T_TMEJ; }
        else if (r.del_size < 4 && r.ins_size < 5) { junction_type_data[1] += r.count; r.type = T_NHEJ; }
        else if (r.del_size >= 4 && r.mh_size < 2 && r.ins_size < 5) { junction_type_data[4] += r.count; r.type = T_NONMH; }
        else if (r.ins_size >= 5) { junction_type_data[2] += r.count; r.type = T_INS; }
        else { junction_type_data[3] += r.count; r.type = T_OTHER; }
    }

    // ---- pass 2 (:396-474): natsorted((count, first-seen rank), reverse=True); label sums over every pattern, rows
    // and plot tuples for the patterns at or above PatternThreshold
    std::vector<int64_t> order((size_t)n);
    for (int64_t k = 0; k < n; ++k) order[(size_t)k] = k;
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        return R[(size_t)a].count != R[(size_t)b].count ? R[(size_t)a].count > R[(size_t)b].count : a > b;
    });
    std::string out;
    int64_t np = 0;
    int n_labels = 0;
    for (int64_t oi = 0; oi < n; ++oi) {
        const int64_t k = order[(size_t)oi];
        const Row &r = R[(size_t)k];
        if (label_order[0] < 0 || std::find(label_order, label_order + n_labels, r.type) == label_order + n_labels)
            label_order[n_labels++] = r.type;
        label_sums[r.type] += (double)r.count / (double)total_scars;
        if (r.key_frequency < in->pattern_threshold) continue;
        if (plot && np < plot_capacity) {
            double *p = plot + np * 6;
            p[0] = r.type; p[1] = r.key_frequency; p[2] = r.lft_del; p[3] = r.rt_del; p[4] = r.mh_size; p[5] = r.ins_size;
        }
        ++np;
        // the row (:300-327,386-394,472-473)
        const scar_record &rec = in->first[k];
        const int rlen = (int)(in->read_off[k + 1] - in->read_off[k]);
        const std::string stored((const char *)in->reads + in->read_off[k], (size_t)rlen);
        std::string insertion = (rec.flags & SCAR_FL_INSERTION) ? pyslice(stored, rec.clj, rec.crj) : std::string();
        std::string microhomology = (rec.flags & SCAR_FL_MICROHOM) ? pyslice(stored, rec.crj, rec.clj) : std::string();
        std::string consensus = stored;
        long a = rec.clj, b = rec.crj, c = rec.tlj, d = rec.trj;
        if (rc) {
            insertion = rcomp(insertion); microhomology = rcomp(microhomology); consensus = rcomp(consensus);
            a = rlen - rec.crj; b = rlen - rec.clj; c = L - rec.trj; d = L - rec.tlj;
        }
        const int32_t *o = in->pattern + k * 16;
        std::string f[8];                                                    // the eight homeology columns
        std::string lft_template, rt_template;
        if (r.type == T_NONMH || r.type == T_INS) {
            if (o[0] >= 0) { f[0] = std::to_string(o[0]); f[1] = std::to_string(o[1]); f[2] = pyslice(consensus, o[2], o[2] + o[3]); f[3] = pyslice(oriented, o[4], o[4] + o[5]); }
            if (o[6] >= 0) { f[4] = std::to_string(o[6]); f[5] = std::to_string(o[7]); f[6] = pyslice(consensus, o[8], o[8] + o[9]); f[7] = pyslice(oriented, o[10], o[10] + o[11]); }
        }
        if (r.type == T_INS) {
            if (o[12] >= 0) { lft_template = pyslice(region, o[12], o[12] + 5); if (rc) lft_template = rcomp(lft_template); }
            if (o[13] >= 0) { rt_template = pyslice(region, o[13], o[13] + 5); if (rc) rt_template = rcomp(rt_template); }
        }
        out += std::to_string(r.count); out += '\t';
        out += py_float_repr((double)r.count / (double)total_scars); out += '\t';
        out += TYPE_NAMES[r.type]; out += '\t';
        out += std::to_string(r.lft_del); out += '\t';
        out += std::to_string(r.rt_del); out += '\t';
        out += std::to_string(r.del_size); out += '\t';
        out += microhomology; out += '\t';
        out += std::to_string(r.mh_size); out += '\t';
        out += insertion; out += '\t';
        out += std::to_string(r.ins_size); out += '\t';
        out += lft_template; out += '\t';
        out += rt_template; out += '\t';
        out += std::to_string(a); out += '\t';
        out += std::to_string(b); out += '\t';
        out += std::to_string(c); out += '\t';
        out += std::to_string(d); out += '\t';
        for (int i = 0; i < 8; ++i) { out += f[i]; out += '\t'; }
        out += consensus; out += '\t';
        out += oriented; out += '\n';
    }
    *n_plot = np;
    *rows_bytes = (int64_t)out.size();
    g_rows.swap(out);                              // handed out by scar_frequency_rows_fetch
    if (plot && np > plot_capacity) return scar_fail(SCAR_ERR_INVALID, "plot buffer too small: %lld rows", (long long)np);
    return SCAR_OK;
}
