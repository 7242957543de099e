// scar_alignment_core.h — AlignmentProcessing.alignment_processing (reference scarmapper/AlignmentProcessing.pyx:18-149)
// for one gapped alignment, host + device.  The reference walks the two gapped strings (consensus vs genome, '-' =
// gap) position by position inside the central 10-90 % of the alignment and counts deleted / inserted columns left
// and right of the cut position; at the first column of every event it also collects a string: the microhomology
// after a deletion, the inserted bases, or a run of mismatches ("snv").  Every such string is a slice of one of the
// two inputs, so the result here is fixed-width: four counters plus (start, length) spans -
//   microhomology : gapped_genomic  [start : start + len]      (AlignmentProcessing.pyx:62-86)
//   insertion     : gapped_consensus[start : start + len]      (:97-107)
//   snv           : gapped_consensus[start : start + len]      (:112-122)
// The 10 % / 90 % limits are the reference's double products 0.10 * gap_genome_len and 0.90 * gap_genome_len compared
// with the integer position (:34-35,50): the same IEEE-754 multiplications and comparisons are done here.
// tests/test_alignment.py pins this header (compiled for the host) to the compiled reference module.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SCAR_ALN_HD __host__ __device__ __forceinline__
#else
#define SCAR_ALN_HD static inline
#endif

struct ScarAlignmentCounts {
    int32_t left_deletion, right_deletion, left_insertion, right_insertion;
    int32_t n_microhomology, n_insertion, n_snv;
    int32_t overflow;          // 1: more events of one kind than max_spans (the extra spans are dropped)
};

// spans: [3][max_spans][2] int32 (kind 0 microhomology, 1 insertion, 2 snv; start, length)
SCAR_ALN_HD void scar_alignment_walk(const uint8_t *genomic, int n_genomic, const uint8_t *consensus, int n_consensus,
                                     int cutposition, int gap_con_len, int gap_genome_len, int max_spans,
                                     ScarAlignmentCounts &out, int32_t *spans)
{
    out.left_deletion = out.right_deletion = out.left_insertion = out.right_insertion = 0;
    out.n_microhomology = out.n_insertion = out.n_snv = 0;
    out.overflow = 0;
    const double lower_limit = 0.10 * (double)gap_genome_len, upper_limit = 0.90 * (double)gap_genome_len;
    bool indel = false;
    const int n = n_genomic < n_consensus ? n_genomic : n_consensus;          // zip() stops at the shorter string
    for (int position = 0; position < n; ++position) {
        const uint8_t c = consensus[position], g = genomic[position];
        if (c == g || (double)position < lower_limit || (double)position > upper_limit) {
            indel = false;
        } else if (c == '-') {                         // deletion: count it, look for microhomology at its first column
            if (position <= cutposition) ++out.left_deletion; else ++out.right_deletion;
            int m = 0;
            if (!indel) {
                int new_position = position;
                while (new_position < gap_con_len && consensus[new_position] == '-') ++new_position;
                int left_step = position, right_step = new_position;
                for (;;) {
                    if (right_step >= gap_con_len || left_step >= new_position || consensus[right_step] == 'N') break;
                    if (genomic[left_step] != consensus[right_step]) break;
                    ++m; ++right_step; ++left_step;
                }
            }
            indel = true;
            if (m) {
                if (out.n_microhomology < max_spans) {
                    spans[(0 * max_spans + out.n_microhomology) * 2] = position;
                    spans[(0 * max_spans + out.n_microhomology) * 2 + 1] = m;
                } else out.overflow = 1;
                ++out.n_microhomology;
            }
        } else if (g == '-' && c != 'N') {             // insertion: count it, capture the inserted run at its first column
            if (position <= cutposition) ++out.left_insertion; else ++out.right_insertion;
            if (!indel) {
                int new_position = position;           // (:100 re-reads len(gapped_genomic) here)
                while (new_position < n_genomic && genomic[new_position] == '-' && consensus[new_position] != 'N') ++new_position;
                if (new_position > position) {
                    if (out.n_insertion < max_spans) {
                        spans[(1 * max_spans + out.n_insertion) * 2] = position;
                        spans[(1 * max_spans + out.n_insertion) * 2 + 1] = new_position - position;
                    } else out.overflow = 1;
                    ++out.n_insertion;
                }
            }
            indel = true;
        } else if (c != g && c != 'N' && g != 'N') {   // mismatch run
            if (!indel) {
                int new_position = position;
                while (new_position < gap_con_len && genomic[new_position] != consensus[new_position] &&
                       genomic[new_position] != '-' && consensus[new_position] != 'N' && consensus[new_position] != '-')
                    ++new_position;
                if (new_position > position) {
                    if (out.n_snv < max_spans) {
                        spans[(2 * max_spans + out.n_snv) * 2] = position;
                        spans[(2 * max_spans + out.n_snv) * 2 + 1] = new_position - position;
                    } else out.overflow = 1;
                    ++out.n_snv;
                }
            }
            indel = true;
        }
    }
}
