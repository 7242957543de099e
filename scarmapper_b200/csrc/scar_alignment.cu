// scar_alignment.cu — AlignmentProcessing.alignment_processing (reference scarmapper/AlignmentProcessing.pyx:18-149)
// over a batch of gapped alignments, one thread per alignment (the walk is sequential by nature: an event's handling
// depends on whether the previous column was part of an event).  The reference module is legacy code (not built by
// scarmapper/setup.py, imported nowhere), named in the north star; the arithmetic lives in scar_alignment_core.h.
#include "scar_internal.cuh"
#include "scar_alignment_core.h"

struct AlnParams {
    scar_ragged genomic, consensus;        // device pointers
    const int32_t *cutposition, *gap_con_len, *gap_genome_len;
    int64_t n;
    int32_t max_spans;
    ScarAlignmentCounts *counts;
    int32_t *spans;
};

__global__ void __launch_bounds__(128) scar_alignment_kernel(const AlnParams P)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t *g, *c; int ng, nc;
        get_item(P.genomic, i, g, ng);
        get_item(P.consensus, i, c, nc);
        ScarAlignmentCounts out;
        scar_alignment_walk(g, ng, c, nc, P.cutposition[i], P.gap_con_len[i], P.gap_genome_len[i], P.max_spans, out,
                            P.spans + (size_t)i * 3 * P.max_spans * 2);
        P.counts[i] = out;
    }
}

cudaError_t scar_launch_alignment(const scar_ragged &genomic, const scar_ragged &consensus, const int32_t *cutposition,
                                  const int32_t *gap_con_len, const int32_t *gap_genome_len, int64_t n, int32_t max_spans,
                                  void *counts, int32_t *spans, int sm_count, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    AlnParams P;
    P.genomic = genomic; P.consensus = consensus; P.cutposition = cutposition; P.gap_con_len = gap_con_len;
    P.gap_genome_len = gap_genome_len; P.n = n; P.max_spans = max_spans;
    P.counts = reinterpret_cast<ScarAlignmentCounts *>(counts); P.spans = spans;
    int64_t blocks = (n + 127) / 128;
    if (blocks > (int64_t)sm_count * 16) blocks = (int64_t)sm_count * 16;
    scar_alignment_kernel<<<(unsigned)blocks, 128, 0, stream>>>(P);
    return cudaGetLastError();
}
