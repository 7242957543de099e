// scar_ctx.cuh — the context behind the opaque scar_ctx of include/scar_b200.h, shared by the translation units that
// implement the C ABI (scar_api.cu: set-up, stage entry points, host-buffer pipeline; scar_ingest.cu: FASTQ files).
#pragma once
#include <algorithm>
#include <vector>

#include "scar_internal.cuh"

int scar_fail(int code, const char *fmt, ...);      // sets scar_last_error(), returns code
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return scar_fail(SCAR_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

static const int NBUF = 3;

struct HostBuf {   // device scratch of one in-flight chunk of scar_process_host
    uint8_t *seq = nullptr, *f0 = nullptr, *f1 = nullptr; size_t seq_cap = 0, f0_cap = 0, f1_cap = 0;
    int64_t *seq_off = nullptr, *f0_off = nullptr, *f1_off = nullptr;
    int32_t *seq_len = nullptr, *f0_len = nullptr, *f1_len = nullptr;
    uint64_t *cells = nullptr; uint32_t *lens = nullptr; uint16_t *samples = nullptr; scar_record *records = nullptr;
    int64_t reads_cap = 0;
    cudaStream_t stream = nullptr; cudaEvent_t done = nullptr;
};

struct scar_ctx {
    int device = 0, sm_count = 148, sm_total = 148, platform = 0, mismatch = 1, n_targets = 0, n_samples = 0;
    int W = 0, max_read_len = 0, min_length = 0, hr_len = 0, do_phasing = 0;
    double n_limit = 0.0; uint64_t hr_code = 0;
    std::vector<ScarTargetMeta> h_targets;
    int32_t *d_sample_target = nullptr;
    unsigned char *d_blob = nullptr; uint32_t blob_bytes = 0, meta_off = 0, phase_off = 0, nmax_off = 0;
    bool smem_tables = true, compact_tables = false;
    uint8_t *d_class_map = nullptr; DemuxString *d_right = nullptr, *d_left = nullptr; uint64_t *d_peq = nullptr;
    uint16_t *d_first_sample = nullptr; int n_right = 0, n_left = 0, n_classes = 0;
    uint8_t *d_fast_codes = nullptr; int demux_fast = 0;
    uint64_t *d_lut_right = nullptr, *d_lut_left = nullptr; int lut_n_right = 0, lut_n_left = 0;
    ScarTableSlot *d_slots = nullptr; uint64_t capacity = 0;
    uint32_t *d_claimed = nullptr;         // indices of the claimed slots, in claim order
    unsigned long long *d_acc = nullptr;   // [n_entries, unassigned, collisions, export_count] + counters + phase hist
    unsigned long long *d_n_entries = nullptr, *d_unassigned = nullptr, *d_collisions = nullptr, *d_export_n = nullptr;
    unsigned long long *d_counters = nullptr, *d_phase_hist = nullptr; size_t acc_words = 0;
    int32_t *d_status = nullptr;
    HostBuf buf[NBUF]; int64_t chunk_reads = 1 << 20;
    bool weak_keys = false;                 // test hook (scar_ctx_set_verify(ctx, 2)): key tags without the segment
    bool verify = false;                    // scar_process_host: one resident chunk + exact key verification
    int fused_ok = -1;                      // -1 unknown, 1: scar_run_dev takes the fused kernel, 0: K1 -> K3 -> K4
    // GPU FASTQ path
    uint8_t *d_text = nullptr; size_t text_cap = 0;
    uint32_t *d_tile_counts = nullptr; size_t tile_counts_cap = 0;
    int64_t *d_tile_offsets = nullptr; size_t tile_offsets_cap = 0;
    int64_t *d_nl = nullptr; size_t nl_cap = 0;
    unsigned long long *d_part_cursor = nullptr;     // region cursors of the peer-memory export
    scar_record *d_fq_records = nullptr; size_t fq_records_cap = 0; int64_t fq_records_n = 0;   // records of the last FASTQ-text call
    std::vector<cudaEvent_t> fq_events;
    // FASTQ file path (scar_ingest.cu): sequence views of the whole run, pinned staging pieces (kept between calls)
    int64_t *d_fq_seq_off = nullptr; size_t fq_seq_off_cap = 0;
    int32_t *d_fq_seq_len = nullptr; size_t fq_seq_len_cap = 0;
    uint8_t *h_piece[4] = {nullptr, nullptr, nullptr, nullptr}; int64_t h_piece_bytes = 0;
    int launches = 0;
    uint8_t *d_gather_small = nullptr, *d_gather_out = nullptr; size_t gather_small_cap = 0, gather_out_cap = 0;   // scar_fastq_gather_host scratch (kept: cudaFree stalls)
};


// (re)allocate a device array of at least `need` elements; the contents are not kept
template <typename T> static inline int ensure(T **p, size_t *cap, size_t need)
{
    if (need <= *cap && *p) return SCAR_OK;
    if (*p) CU(cudaFree(*p));
    *p = nullptr;
    size_t n = std::max<size_t>(need + need / 8, 256);
    if (cudaError_t e = cudaMalloc((void **)p, n * sizeof(T)); e != cudaSuccess) {
        *p = nullptr; *cap = 0;
        cudaGetLastError();
        return scar_fail(e == cudaErrorMemoryAllocation ? SCAR_ERR_NOMEM : SCAR_ERR_CUDA, "no device memory for %zu bytes: %s",
                         n * sizeof(T), cudaGetErrorString(e));
    }
    *cap = n;
    return SCAR_OK;
}


// grow a device array and keep its first `keep` elements
template <typename T> static inline int ensure_keep(T **p, size_t *cap, size_t need, size_t keep)
{
    if (need <= *cap && *p) return SCAR_OK;
    T *q = nullptr;
    const size_t n = std::max<size_t>(need + need / 4, 256);
    if (cudaError_t e = cudaMalloc((void **)&q, n * sizeof(T)); e != cudaSuccess) {
        cudaGetLastError();
        return scar_fail(e == cudaErrorMemoryAllocation ? SCAR_ERR_NOMEM : SCAR_ERR_CUDA, "no device memory for %zu bytes: %s",
                         n * sizeof(T), cudaGetErrorString(e));
    }
    if (*p && keep) CU(cudaMemcpy(q, *p, keep * sizeof(T), cudaMemcpyDeviceToDevice));
    if (*p) CU(cudaFree(*p));
    *p = q; *cap = n;
    return SCAR_OK;
}


int scar_ensure_hostbuf(scar_ctx *c, HostBuf &b, int64_t reads);   // streams + per-chunk device scratch for `reads` reads
int scar_check_status(scar_ctx *c);                                // device status word -> error code
bool scar_fused_path(scar_ctx *c);                                 // scar_run_dev takes the fused kernel for this context
int scar_grow_read_len(scar_ctx *c, int max_len);                  // re-plan for longer reads (streams must be idle)
