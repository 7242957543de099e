// scar_fastq_dev.cu — FASTQ record indexing ON THE GPU (the loader's front end for text already in HBM).
//
// Same contract as the host indexer scar_fastq_index (scar_fastq.cu) — FASTQ_Reader.seq_read
// (reference Valkyries/FASTQ_Tools.py:623-653) + the header split of index_matching
// (scarmapper/INDEL_Processing.py:1155-1156) — but over a text buffer resident in device memory:
//   pass 1  count '\n' per 4 KB tile (16-byte vector loads, popcount of the byte-equality mask)
//   pass 2  exclusive scan of the tile counts (two levels: 1024 tiles per CTA, then one CTA over the CTA sums)
//   pass 3  every tile writes the positions of its newlines at its scanned offset
//   pass 4  one thread per record: lines 4r .. 4r+3 -> (offset, length) views of the sequence and of
//           name.split(":")[-1].split("+")[0..1]; whitespace stripped as str.strip() does; a
//           sequence/quality length mismatch or a header without '+' raises the error flags the host
//           indexer returns as SCAR_ERR_INVALID.
// The views point into the text buffer itself, so K1/K2/K4 read the fields in place.
#include "scar_internal.cuh"

#define FQ_TILE 4096           // bytes per CTA tile (256 threads x 16 bytes)

__device__ __forceinline__ uint32_t newline_mask16(const uint4 v)
{
    // bit k set iff byte k of the 16-byte vector is '\n'
    uint32_t m = 0;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t z = zero_bytes(w[i] ^ 0x0A0A0A0Au) >> 7;           // bit 0 of each byte
        m |= ((z | (z >> 7) | (z >> 14) | (z >> 21)) & 0xFu) << (4 * i);
    }
    return m;
}

// 16 bytes at tile offset; bytes past n_bytes read as 0
__device__ __forceinline__ uint4 load16(const uint8_t *text, int64_t off, int64_t n_bytes)
{
    if (off + 16 <= n_bytes) return __ldg(reinterpret_cast<const uint4 *>(text + off));   // text is 16-byte aligned
    uint4 v = make_uint4(0, 0, 0, 0);
    uint8_t *b = reinterpret_cast<uint8_t *>(&v);
    for (int k = 0; k < 16 && off + k < n_bytes; ++k) b[k] = text[off + k];
    return v;
}

__global__ void __launch_bounds__(256) fq_count_kernel(const uint8_t *text, int64_t n_bytes, uint32_t *tile_counts)
{
    const int64_t off = (int64_t)blockIdx.x * FQ_TILE + threadIdx.x * 16;
    int c = off < n_bytes ? __popc(newline_mask16(load16(text, off, n_bytes))) : 0;
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    __shared__ int s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += s[i];
        tile_counts[blockIdx.x] = (uint32_t)t;
    }
}

// Two-level exclusive scan of the tile counts.  Level 1: every CTA scans 1024 tile counts (local prefix into
// tile_offsets, the CTA's sum into block_sums).  Level 2: ONE CTA scans the block sums in place (a 3.5 GB
// text has 845 K tiles = 826 block sums) and writes the grand total.  A tile's offset is then
// block_sums[tile / 1024] + tile_offsets[tile].
#define FQ_SCAN_BLOCK 1024

__device__ __forceinline__ unsigned long long cta_exclusive_scan(unsigned long long v, unsigned long long *s_warp,
                                                                 unsigned long long &cta_total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, w, d);
            if (lane >= d) w += y;
        }
        s_warp[lane] = w;                             // inclusive over warps
    }
    __syncthreads();
    cta_total = s_warp[31];
    const unsigned long long before = (warp ? s_warp[warp - 1] : 0ull) + (x - v);
    __syncthreads();                                  // s_warp is reused by the caller's next round
    return before;
}

__global__ void __launch_bounds__(FQ_SCAN_BLOCK) fq_scan1_kernel(const uint32_t *tile_counts, int64_t n_tiles,
                                                                 int64_t *tile_offsets, unsigned long long *block_sums)
{
    __shared__ unsigned long long s_warp[32];
    const int64_t i = (int64_t)blockIdx.x * FQ_SCAN_BLOCK + threadIdx.x;
    unsigned long long total;
    const unsigned long long before = cta_exclusive_scan(i < n_tiles ? tile_counts[i] : 0ull, s_warp, total);
    if (i < n_tiles) tile_offsets[i] = (int64_t)before;
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(FQ_SCAN_BLOCK) fq_scan2_kernel(unsigned long long *block_sums, int64_t n_blocks,
                                                                 unsigned long long *total)
{
    __shared__ unsigned long long s_warp[32];
    unsigned long long carry = 0;
    for (int64_t base = 0; base < n_blocks; base += FQ_SCAN_BLOCK) {
        const int64_t i = base + threadIdx.x;
        unsigned long long round_total;
        const unsigned long long before = cta_exclusive_scan(i < n_blocks ? block_sums[i] : 0ull, s_warp, round_total);
        if (i < n_blocks) block_sums[i] = carry + before;
        carry += round_total;
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256) fq_positions_kernel(const uint8_t *text, int64_t n_bytes, const int64_t *tile_offsets,
                                                           const unsigned long long *block_sums, int64_t *nl_pos,
                                                           int64_t nl_capacity, int64_t byte_base)
{
    const int64_t off = (int64_t)blockIdx.x * FQ_TILE + threadIdx.x * 16;
    const uint32_t m = off < n_bytes ? newline_mask16(load16(text, off, n_bytes)) : 0u;
    const int c = __popc(m);
    // exclusive prefix of c inside the CTA
    int x = c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= d) x += y;
    }
    __shared__ int s[8];
    if (lane == 31) s[warp] = x;
    __syncthreads();
    int before = x - c;
    for (int w = 0; w < warp; ++w) before += s[w];
    int64_t dst = (int64_t)block_sums[blockIdx.x / FQ_SCAN_BLOCK] + tile_offsets[blockIdx.x] + before;
    for (uint32_t mm = m; mm; mm &= mm - 1) {
        if (dst < nl_capacity) nl_pos[dst] = byte_base + off + (__ffs(mm) - 1);
        ++dst;
    }
}

__device__ __forceinline__ bool fq_space(uint8_t c)
{
    return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f' || c == '\n';
}

struct FqViews {
    int64_t *seq_off; int32_t *seq_len; int64_t *f0_off; int32_t *f0_len; int64_t *f1_off; int32_t *f1_len;
};

// status bits: 8 = sequence/quality length mismatch, 16 = header without '+'
// records first_record .. first_record + n_records - 1 of the text; view r - first_record is written
__global__ void __launch_bounds__(256) fq_records_kernel(const uint8_t *text, int64_t n_bytes, const int64_t *nl_pos,
                                                         int64_t n_newlines, int64_t first_record, int64_t n_records,
                                                         int want_fields, FqViews V, int32_t *status, int32_t *max_len)
{
    int longest = 0;                    // longest sequence line this thread has seen (one atomic per warp at the end)
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_records;
         r += (int64_t)gridDim.x * blockDim.x) {
        int64_t ls[4], le[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = 4 * (first_record + r) + k;
            ls[k] = i == 0 ? 0 : nl_pos[i - 1] + 1;
            le[k] = i < n_newlines ? nl_pos[i] : n_bytes;          // last record without a final newline
        }
        int64_t s0 = ls[1], s1 = le[1], q0 = ls[3], q1 = le[3];
        while (s0 < s1 && fq_space(text[s0])) ++s0;
        while (s1 > s0 && fq_space(text[s1 - 1])) --s1;
        while (q0 < q1 && fq_space(text[q0])) ++q0;
        while (q1 > q0 && fq_space(text[q1 - 1])) --q1;
        if (s1 - s0 != q1 - q0) atomicOr(status, 8);
        V.seq_off[r] = s0; V.seq_len[r] = (int32_t)(s1 - s0);
        longest = max(longest, (int)(s1 - s0));
        if (want_fields) {
            int64_t h0 = ls[0], h1 = le[0];
            if (h1 > h0 && text[h1 - 1] == '\r') --h1;             // text mode: "\r\n" is one newline
            while (h0 < h1 && text[h0] == '@') ++h0;               // .strip("@")
            while (h1 > h0 && text[h1 - 1] == '@') --h1;
            int64_t c = h1;                                        // text after the last ':'
            while (c > h0 && text[c - 1] != ':') --c;
            int64_t p1 = c;
            while (p1 < h1 && text[p1] != '+') ++p1;
            if (p1 >= h1) { atomicOr(status, 16); p1 = h1; }
            int64_t p2 = p1 + 1;
            while (p2 < h1 && text[p2] != '+') ++p2;
            if (p2 > h1) p2 = h1;
            V.f0_off[r] = c; V.f0_len[r] = (int32_t)(p1 - c);
            V.f1_off[r] = p1 < h1 ? p1 + 1 : h1; V.f1_len[r] = p1 < h1 ? (int32_t)(p2 - p1 - 1) : 0;
        }
    }
    if (max_len) {
        longest = __reduce_max_sync(0xFFFFFFFFu, longest);
        if ((threadIdx.x & 31) == 0 && longest > 0) atomicMax(max_len, longest);
    }
}

// host orchestration -----------------------------------------------------------------------------------
int scar_fail(int code, const char *fmt, ...);

// elements (int64) the tile_offsets buffer needs: one per tile + one per 1024 tiles (block sums)
int64_t scar_fastq_offsets_elems(int64_t n_bytes)
{
    const int64_t n_tiles = (n_bytes + FQ_TILE - 1) / FQ_TILE;
    return n_tiles + (n_tiles + FQ_SCAN_BLOCK - 1) / FQ_SCAN_BLOCK + 1;
}

cudaError_t scar_fastq_index_launch(const uint8_t *text, int64_t n_bytes, uint32_t *tile_counts, int64_t *tile_offsets,
                                    unsigned long long *total, cudaStream_t s)
{
    const int64_t n_tiles = (n_bytes + FQ_TILE - 1) / FQ_TILE;
    if (n_tiles == 0) return cudaMemsetAsync(total, 0, sizeof(unsigned long long), s);
    const int64_t n_blocks = (n_tiles + FQ_SCAN_BLOCK - 1) / FQ_SCAN_BLOCK;
    unsigned long long *block_sums = reinterpret_cast<unsigned long long *>(tile_offsets + n_tiles);   // see scar_fastq_offsets_elems
    fq_count_kernel<<<(unsigned)n_tiles, 256, 0, s>>>(text, n_bytes, tile_counts);
    fq_scan1_kernel<<<(unsigned)n_blocks, FQ_SCAN_BLOCK, 0, s>>>(tile_counts, n_tiles, tile_offsets, block_sums);
    fq_scan2_kernel<<<1, FQ_SCAN_BLOCK, 0, s>>>(block_sums, n_blocks, total);
    return cudaGetLastError();
}

// `text` may be a piece of a larger buffer: byte_base is added to every position written
cudaError_t scar_fastq_positions_launch(const uint8_t *text, int64_t n_bytes, const int64_t *tile_offsets, int64_t *nl_pos,
                                        int64_t nl_capacity, int64_t byte_base, cudaStream_t s)
{
    const int64_t n_tiles = (n_bytes + FQ_TILE - 1) / FQ_TILE;
    if (n_tiles == 0) return cudaSuccess;
    const unsigned long long *block_sums = reinterpret_cast<const unsigned long long *>(tile_offsets + n_tiles);
    fq_positions_kernel<<<(unsigned)n_tiles, 256, 0, s>>>(text, n_bytes, tile_offsets, block_sums, nl_pos, nl_capacity, byte_base);
    return cudaGetLastError();
}

cudaError_t scar_fastq_records_launch(const uint8_t *text, int64_t n_bytes, const int64_t *nl_pos, int64_t n_newlines,
                                      int64_t first_record, int64_t n_records, int want_fields, int64_t *seq_off, int32_t *seq_len,
                                      int64_t *f0_off, int32_t *f0_len, int64_t *f1_off, int32_t *f1_len, int32_t *status,
                                      int sm_count, cudaStream_t s, int32_t *max_len)
{
    if (n_records <= 0) return cudaSuccess;
    FqViews V{seq_off, seq_len, f0_off, f0_len, f1_off, f1_len};
    int64_t blocks = (n_records + 255) / 256;
    const int64_t cap = (int64_t)sm_count * 16;
    if (blocks > cap) blocks = cap;
    fq_records_kernel<<<(unsigned)blocks, 256, 0, s>>>(text, n_bytes, nl_pos, n_newlines, first_record, n_records, want_fields, V, status, max_len);
    return cudaGetLastError();
}
