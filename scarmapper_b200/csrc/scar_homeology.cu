// scar_homeology.cu — the per-unique-scar searches of ScarSearch.frequency_output as ONE batched kernel:
// homeology_search (reference scarmapper/INDEL_Processing.py:550-657) and templated_insertion_search
// (:659-700) for every key of a sample's scar table at once, one thread per key.  The arithmetic lives in
// scar_homeology_core.h (host + device; tests/test_homeology_host.py checks it against the reference's own
// methods).  The reference calls these once per scar pattern from a Python loop (10 Levenshtein calls on
// 5-mers and up to 100 slice comparisons per key); with ~1 M patterns per run that loop is the wall-time
// floor of a whole job.
#include "scar_internal.cuh"
#include "scar_homeology_core.h"

struct HomParams {
    scar_ragged cons, ins, targets, regions;       // device pointers
    const int32_t *clj, *crj, *tlj, *trj, *target_idx;
    int32_t n_targets;
    int64_t n;
    ScarHomeology *out;
};

__global__ void __launch_bounds__(128) scar_homeology_kernel(const HomParams P)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t *cons, *ins, *target, *region;
        int n_cons, n_ins, n_target, n_region;
        get_item(P.cons, i, cons, n_cons);
        get_item(P.ins, i, ins, n_ins);
        const int t = P.target_idx[i];
        ScarHomeology h;
        if (t < 0 || t >= P.n_targets) {           // no target: nothing can be found
            h.l_dist = h.r_dist = -1; h.lt_start = h.rt_start = -1;
            h.l_pos = h.l_q_start = h.l_q_len = h.l_s_start = h.l_s_len = 0;
            h.r_pos = h.r_q_start = h.r_q_len = h.r_s_start = h.r_s_len = 0;
        } else {
            get_item(P.targets, t, target, n_target);
            get_item(P.regions, t, region, n_region);
            hom_homeology(cons, n_cons, P.clj[i], P.crj[i], target, n_target, n_ins, P.tlj[i], P.trj[i], h);
            h.lt_start = h.rt_start = -1;
            if (n_ins >= 5) hom_templates(ins, n_ins, region, n_region, P.tlj[i], P.trj[i], h);
        }
        h.pad0 = h.pad1 = 0;
        P.out[i] = h;
    }
}

cudaError_t scar_launch_homeology(const scar_ragged &cons, const scar_ragged &ins, const scar_ragged &targets,
                                  const scar_ragged &regions, const int32_t *clj, const int32_t *crj, const int32_t *tlj,
                                  const int32_t *trj, const int32_t *target_idx, int32_t n_targets, int64_t n, void *out,
                                  cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    HomParams P;
    P.cons = cons; P.ins = ins; P.targets = targets; P.regions = regions;
    P.clj = clj; P.crj = crj; P.tlj = tlj; P.trj = trj; P.target_idx = target_idx; P.n_targets = n_targets; P.n = n;
    P.out = reinterpret_cast<ScarHomeology *>(out);
    int64_t blocks = (n + 127) / 128;
    int dev = 0, sms = 148;                              // grid: at most 16 CTAs of 128 threads per SM of this device
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
    scar_homeology_kernel<<<(unsigned)blocks, 128, 0, stream>>>(P);
    return cudaGetLastError();
}
