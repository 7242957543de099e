// scar_ingest.cu — a whole FASTQ FILE through the path: scar_process_fastq_file.
//
// Replaces, for the batched drop-in, FASTQ_Reader.__init__ / seq_read (reference Valkyries/FASTQ_Tools.py:562-653:
// the file is opened as text or through gzip and parsed one record per call) as the loader in front of
// consensus_demultiplex (scarmapper/INDEL_Processing.py:891-1024).  Here the file is streamed:
//
//   producer thread   plain text : the file is mapped, pieces are copied into pinned staging buffers by a few threads
//                     gzip       : one DEFLATE stream (or several members back to back) inflated by zlib straight
//                                  into the pinned piece - a single stream cannot be split, so it is overlapped
//                     BGZF       : every block names its compressed size (gzip extra field 'BC') and its inflated
//                                  size (trailer): the blocks of a piece are inflated side by side by the threads
//   calling thread    each piece goes to the GPU as it arrives (one H2D copy), is indexed THERE (newline count /
//                     scan / positions / one thread per record, scar_fastq_dev.cu) and the records that have become
//                     complete run through the whole path (K2 demux + fused pack / window / aggregate, or K1-K3-K4).
//
// Nothing about the run has to be known up front: the device text buffer and the record arrays grow, the context is
// re-planned when a longer read shows up (scar_grow_read_len) and the scar table is rebuilt at a larger size before
// it could pass half full (scar_table_reserve).  With scar_ctx_set_verify the whole run - the text stays resident -
// is checked key by key afterwards (exact-or-fail scar-pattern identity).
#include <atomic>
#include <cstdint>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "scar_ctx.cuh"
#include "scar_pgzip.h"
#include "scar_comp.h"

namespace {

double now_s()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

// a few persistent helper threads: run(n, f) calls f(0) .. f(n-1), the caller takes part, returns when all are done
class Workers {
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, cv_done_;
    const std::function<void(int)> *fn_ = nullptr;
    int n_ = 0, next_ = 0, done_ = 0;
    bool quit_ = false;

    void work(std::unique_lock<std::mutex> &lk)
    {
        while (fn_ && next_ < n_) {
            const int i = next_++;
            const std::function<void(int)> *f = fn_;
            lk.unlock();
            (*f)(i);
            lk.lock();
            if (++done_ == n_) { fn_ = nullptr; cv_done_.notify_all(); }
        }
    }

public:
    explicit Workers(int helpers)
    {
        for (int i = 0; i < helpers; ++i)
            th_.emplace_back([this] {
                std::unique_lock<std::mutex> lk(m_);
                for (;;) {
                    cv_.wait(lk, [this] { return quit_ || (fn_ && next_ < n_); });
                    if (quit_) return;
                    work(lk);
                }
            });
    }
    ~Workers()
    {
        { std::lock_guard<std::mutex> lk(m_); quit_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return (int)th_.size() + 1; }
    void run(int n, const std::function<void(int)> &f)
    {
        if (n <= 0) return;
        std::unique_lock<std::mutex> lk(m_);
        fn_ = &f; n_ = n; next_ = 0; done_ = 0;
        cv_.notify_all();
        work(lk);
        cv_done_.wait(lk, [this] { return fn_ == nullptr; });
    }
};

void parallel_copy(Workers &w, uint8_t *dst, const uint8_t *src, int64_t n)
{
    const int64_t seg = 4 << 20;
    const int parts = (int)((n + seg - 1) / seg);
    w.run(parts, [&](int i) { memcpy(dst + (int64_t)i * seg, src + (int64_t)i * seg, (size_t)std::min(seg, n - (int64_t)i * seg)); });
}

struct MappedFile {
    const uint8_t *p = nullptr; int64_t n = 0; int fd = -1;
    ~MappedFile() { reset(); }
    void reset()
    {
        if (p && n) munmap((void *)p, (size_t)n);
        if (fd >= 0) close(fd);
        p = nullptr; n = 0; fd = -1;
    }
    bool open_ro(const char *path, std::string &err)
    {
        fd = open(path, O_RDONLY);
        if (fd < 0) { err = std::string("cannot open ") + path; return false; }
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { err = std::string(path) + " is not a regular file"; return false; }
        n = (int64_t)st.st_size;
        if (n == 0) return true;
        void *m = mmap(nullptr, (size_t)n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { p = nullptr; err = std::string("cannot map ") + path; return false; }
        p = (const uint8_t *)m;
        madvise(m, (size_t)n, MADV_SEQUENTIAL);
        return true;
    }
};

// what the library hands back with a run and frees in scar_fastq_run_release
struct RunOwner {
    MappedFile text_map;                 // plain text: the host text IS the mapping
    uint8_t *text_heap = nullptr;        // gzip: inflated text
    int64_t *seq_off = nullptr; int32_t *seq_len = nullptr; scar_record *records = nullptr;
    ~RunOwner() { free(text_heap); free(seq_off); free(seq_len); free(records); }
};

struct BgzfBlock { int64_t in_off; int32_t in_len, out_len; uint32_t crc; };

// BGZF (SAM/BAM spec 4.1): every member is a gzip member with FEXTRA and a 'B','C' subfield holding the member's total
// size - 1.  Returns false when the file is not BGZF from the first byte to the last.
bool scan_bgzf(const uint8_t *p, int64_t n, std::vector<BgzfBlock> &blocks)
{
    blocks.clear();
    int64_t pos = 0;
    while (pos < n) {
        if (pos + 18 > n || p[pos] != 0x1f || p[pos + 1] != 0x8b || p[pos + 2] != 8 || !(p[pos + 3] & 4)) return false;
        if (p[pos + 3] & ~4) return false;                        // FNAME / FCOMMENT / FHCRC: not written by BGZF tools
        const int xlen = p[pos + 10] | (p[pos + 11] << 8);
        if (pos + 12 + xlen > n) return false;
        int bsize = -1;
        for (int q = 0; q + 4 <= xlen;) {
            const uint8_t *f = p + pos + 12 + q;
            const int slen = f[2] | (f[3] << 8);
            if (f[0] == 'B' && f[1] == 'C' && slen == 2 && q + 6 <= xlen) bsize = (f[4] | (f[5] << 8)) + 1;
            q += 4 + slen;
        }
        if (bsize < 12 + xlen + 8 || pos + bsize > n) return false;
        const uint8_t *t = p + pos + bsize - 8;
        BgzfBlock b;
        b.in_off = pos + 12 + xlen; b.in_len = bsize - 12 - xlen - 8;
        b.crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        b.out_len = (int32_t)((uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24));
        if (b.out_len < 0 || b.out_len > (1 << 16)) return false;
        blocks.push_back(b);
        pos += bsize;
    }
    return !blocks.empty();
}

// ---- producer: fills pinned pieces in file order -------------------------------------------------------------------
struct Piece { uint8_t *buf = nullptr; int64_t len = 0; bool last = false; };

struct Stream {
    // ring of pieces between the producer thread and the calling thread
    std::mutex m; std::condition_variable cv;
    Piece slot[4]; int n_slots = 3;
    int64_t produced = 0, consumed = 0;          // pieces
    bool stop = false;                            // the consumer has all it wants
    std::string error;                            // set by the producer; ends the stream
    int64_t cap = 0;                              // bytes per piece
    uint8_t **pinned = nullptr; int64_t pinned_bytes = 0; int device = 0;   // the context's cached buffers (>= cap each)

    Piece *acquire_free()                         // producer
    {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return stop || produced - consumed < n_slots; });
        if (stop) return nullptr;
        const int i = (int)(produced % n_slots);
        lk.unlock();
        if (!pinned[i]) {                         // first use of this slot
            cudaSetDevice(device);
            if (cudaHostAlloc((void **)&pinned[i], (size_t)pinned_bytes, cudaHostAllocDefault) != cudaSuccess) {
                pinned[i] = nullptr;
                cudaGetLastError();
                fail("no pinned host memory for a staging piece");
                return nullptr;
            }
        }
        slot[i].buf = pinned[i];
        return &slot[i];
    }
    void publish() { { std::lock_guard<std::mutex> lk(m); ++produced; } cv.notify_all(); }
    void fail(const std::string &e)
    {
        { std::lock_guard<std::mutex> lk(m); if (error.empty()) { error = e; slot[produced % n_slots].len = 0; slot[produced % n_slots].last = true; ++produced; } }
        cv.notify_all();
    }
    Piece *wait_filled()                          // consumer
    {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return produced > consumed; });
        return &slot[consumed % n_slots];
    }
    void release() { { std::lock_guard<std::mutex> lk(m); ++consumed; } cv.notify_all(); }
    void request_stop() { { std::lock_guard<std::mutex> lk(m); stop = true; } cv.notify_all(); }
};

struct HostText {                                 // inflated text kept for the caller (gzip sources)
    uint8_t *p = nullptr; int64_t n = 0, cap = 0;
    ~HostText() { free(p); }
    bool append(Workers &w, const uint8_t *src, int64_t len)
    {
        if (n + len > cap) {
            const int64_t want = std::max<int64_t>(n + len, cap + cap / 2 + (64 << 20));
            uint8_t *q = (uint8_t *)realloc(p, (size_t)want);     // (large blocks: mremap, no copy)
            if (!q) return false;
            p = q; cap = want;
        }
        parallel_copy(w, p + n, src, len);
        n += len;
        return true;
    }
};

void produce_plain(Stream &S, Workers &W, const MappedFile &F)
{
    int64_t off = 0;
    do {
        Piece *pc = S.acquire_free();
        if (!pc) return;
        pc->len = std::min(S.cap, F.n - off);
        parallel_copy(W, pc->buf, F.p + off, pc->len);
        off += pc->len;
        pc->last = off >= F.n;
        S.publish();
    } while (off < F.n);
}

void produce_gzip(Stream &S, Workers &W, const MappedFile &F, HostText *keep)
{
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 15 + 32) != Z_OK) { S.fail("zlib: inflateInit2 failed"); return; }
    int64_t in_off = 0;
    bool finished = false, member_open = false;
    while (!finished) {
        Piece *pc = S.acquire_free();
        if (!pc) break;
        zs.next_out = pc->buf;
        int64_t out_left = S.cap;
        while (out_left > 0 && !finished) {
            if (zs.avail_in == 0) {
                if (in_off >= F.n) {
                    if (member_open) { inflateEnd(&zs); S.fail("gzip: unexpected end of file"); return; }
                    finished = true;
                    break;
                }
                const int64_t take = std::min<int64_t>(F.n - in_off, 1 << 30);
                zs.next_in = const_cast<Bytef *>(F.p + in_off); zs.avail_in = (uInt)take;
                in_off += take;
            }
            if (!member_open) {
                // between members: zero padding is skipped (as gzip and Python's gzip module do); anything else must
                // be the next member's header
                while (zs.avail_in && *zs.next_in == 0) { ++zs.next_in; --zs.avail_in; }
                if (!zs.avail_in) continue;
                member_open = true;
            }
            zs.avail_out = (uInt)std::min<int64_t>(out_left, 1 << 30);
            const uInt before = zs.avail_out;
            const int ret = inflate(&zs, Z_NO_FLUSH);
            out_left -= (int64_t)(before - zs.avail_out);
            if (ret == Z_STREAM_END) {
                member_open = false;
                inflateReset(&zs);                                // the next member (if any) starts with its own header
            } else if (ret != Z_OK && ret != Z_BUF_ERROR) {
                const std::string msg = std::string("gzip: ") + (zs.msg ? zs.msg : "corrupt stream");
                inflateEnd(&zs);
                S.fail(msg);
                return;
            }
        }
        pc->len = S.cap - out_left;
        pc->last = finished;
        if (keep && pc->len && !keep->append(W, pc->buf, pc->len)) { inflateEnd(&zs); S.fail("out of host memory for the inflated text"); return; }
        S.publish();
    }
    inflateEnd(&zs);
}

// one gzip stream (or several members) inflated by n_threads threads: two-pass speculative decoding, scar_pgzip.h
void produce_gzip_parallel(Stream &S, Workers &W, const MappedFile &F, HostText *keep, int n_threads)
{
    scar_pgzip::ParallelGunzip pg(F.p, F.n, n_threads, 0);
    bool finished = false;
    while (!finished) {
        Piece *pc = S.acquire_free();
        if (!pc) break;
        int64_t got = 0;
        while (got < S.cap) {
            const int64_t r = pg.read(pc->buf + got, S.cap - got);
            if (r < 0) { S.fail(pg.error()); return; }
            if (r == 0) { finished = true; break; }
            got += r;
        }
        pc->len = got;
        pc->last = finished;
        if (keep && got && !keep->append(W, pc->buf, got)) { S.fail("out of host memory for the inflated text"); return; }
        S.publish();
    }
}

void produce_bgzf(Stream &S, Workers &W, const MappedFile &F, const std::vector<BgzfBlock> &blocks, HostText *keep)
{
    size_t b0 = 0;
    std::vector<int64_t> out_off;
    std::atomic<int> bad(0);
    while (b0 < blocks.size()) {
        Piece *pc = S.acquire_free();
        if (!pc) return;
        // the blocks of this piece and where each one lands in it
        out_off.clear();
        int64_t total = 0;
        size_t b1 = b0;
        while (b1 < blocks.size() && total + blocks[b1].out_len <= S.cap) { out_off.push_back(total); total += blocks[b1].out_len; ++b1; }
        const int n = (int)(b1 - b0);
        const int per = std::max(1, (n + 4 * W.size() - 1) / (4 * W.size()));      // blocks per task
        W.run((n + per - 1) / per, [&](int task) {
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
            for (int k = task * per; k < std::min(n, (task + 1) * per); ++k) {
                const BgzfBlock &b = blocks[b0 + (size_t)k];
                if (b.out_len == 0) continue;                     // (the empty end-of-file block)
                inflateReset(&zs);
                zs.next_in = const_cast<Bytef *>(F.p + b.in_off); zs.avail_in = (uInt)b.in_len;
                zs.next_out = pc->buf + out_off[(size_t)k]; zs.avail_out = (uInt)b.out_len;
                const int ret = inflate(&zs, Z_FINISH);
                if (ret != Z_STREAM_END || zs.avail_out != 0 ||
                    (uint32_t)crc32(crc32(0L, Z_NULL, 0), pc->buf + out_off[(size_t)k], (uInt)b.out_len) != b.crc) bad = 1;
            }
            inflateEnd(&zs);
        });
        if (bad) { S.fail("BGZF: a block does not inflate to its recorded size / CRC"); return; }
        pc->len = total;
        b0 = b1;
        pc->last = b0 >= blocks.size();
        if (keep && total && !keep->append(W, pc->buf, total)) { S.fail("out of host memory for the inflated text"); return; }
        S.publish();
    }
}

}   // namespace

// ---- stored sequences of chosen reads, gathered from the text that is resident on the device ----------------------
namespace {
__constant__ uint8_t GATHER_COMP[256];

__device__ __forceinline__ void gather_span(int platform, int64_t off, int32_t rawlen, int64_t &a, int32_t &len)
{
    if (platform == SCAR_PLATFORM_TRUSEQ) { a = off + (rawlen > 6 ? 6 : rawlen); len = rawlen > 12 ? rawlen - 12 : 0; }
    else { a = off; len = rawlen; }
}

__global__ void gather_lengths_kernel(const int64_t *rows, int64_t n, const int32_t *seq_len, int platform, int32_t *out_len)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t rawlen = seq_len[rows[k]];
        out_len[k] = platform == SCAR_PLATFORM_TRUSEQ ? (rawlen > 12 ? rawlen - 12 : 0) : rawlen;
    }
}

// one warp per row: coalesced stores; a reversed row reads its source backwards through the complement table
__global__ void gather_rows_kernel(const uint8_t *text, const int64_t *seq_off, const int32_t *seq_len, const int64_t *rows,
                                   const uint8_t *flip, const int64_t *out_off, int64_t n, int platform, uint8_t *out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp; k < n; k += n_warps) {
        const int64_t r = rows[k];
        int64_t a; int32_t len;
        gather_span(platform, seq_off[r], seq_len[r], a, len);
        const bool rev = (platform == SCAR_PLATFORM_RAMSDEN) != (flip && flip[k]);
        uint8_t *dst = out + out_off[k];
        const uint8_t *src = text + a;
        if (rev) for (int i = lane; i < len; i += 32) dst[i] = GATHER_COMP[src[len - 1 - i]];
        else for (int i = lane; i < len; i += 32) dst[i] = src[i];
    }
}
}   // namespace

extern "C" int scar_fastq_gather_host(scar_ctx *c, const int64_t *rows, int64_t n_rows, const uint8_t *flip, uint8_t *out,
                                      int64_t out_capacity, int64_t *out_off)
{
    if (!c || (!rows && n_rows) || !out_off || n_rows < 0) return scar_fail(SCAR_ERR_INVALID, "bad argument");
    out_off[0] = 0;
    if (n_rows == 0) return SCAR_OK;
    if (!c->d_text || !c->d_fq_seq_off || !c->d_fq_seq_len || c->fq_records_n <= 0)
        return scar_fail(SCAR_ERR_INVALID, "no FASTQ text is resident on the device (scar_process_fastq_file first)");
    for (int64_t k = 0; k < n_rows; ++k)
        if (rows[k] < 0 || rows[k] >= c->fq_records_n) return scar_fail(SCAR_ERR_INVALID, "row %lld names read %lld of %lld", (long long)k, (long long)rows[k], (long long)c->fq_records_n);
    CU(cudaSetDevice(c->device));
    const bool trace = getenv("SCAR_TRACE") != nullptr;
    const double t_begin = now_s();
    double t_prev = t_begin;
    std::string timeline;
    auto mark = [&](const char *what) {
        if (!trace) return;
        const double t = now_s();
        char b[96];
        snprintf(b, sizeof b, " %s %.1f ms,", what, 1e3 * (t - t_prev));
        timeline += b; t_prev = t;
    };
    static bool table_ready[64] = {false};
    if (!table_ready[c->device & 63]) {
        uint8_t t[256];
        for (int i = 0; i < 256; ++i) t[i] = scar_comp((uint8_t)i);
        CU(cudaMemcpyToSymbol(GATHER_COMP, t, 256));
        table_ready[c->device & 63] = true;
    }
    // (allocations are the expensive part of this call and cudaFree stalls unpredictably: two scratch blocks kept in
    // the context, one for the small arrays, one for the bytes)
    struct Part { void *p; } d_small, d_out, d_rows, d_off, d_len, d_flip;
    const size_t n = (size_t)n_rows;
    int rc = ensure(&c->d_gather_small, &c->gather_small_cap, n * 21 + 64);
    if (rc) return rc;
    d_small.p = c->d_gather_small; d_out.p = nullptr;
    d_rows.p = d_small.p; d_off.p = (uint8_t *)d_small.p + n * 8; d_len.p = (uint8_t *)d_small.p + n * 16;
    d_flip.p = flip ? (uint8_t *)d_small.p + n * 20 : nullptr;
    mark("table + small buffers");
    cudaStream_t s = c->buf[0].stream;
    CU(cudaMemcpyAsync(d_rows.p, rows, n * 8, cudaMemcpyHostToDevice, s));
    if (flip) CU(cudaMemcpyAsync(d_flip.p, flip, n, cudaMemcpyHostToDevice, s));
    const int grid = c->sm_count * 8;
    gather_lengths_kernel<<<grid, 256, 0, s>>>((const int64_t *)d_rows.p, n_rows, c->d_fq_seq_len, c->platform, (int32_t *)d_len.p);
    CU(cudaGetLastError());
    std::vector<int32_t> len(n);
    CU(cudaMemcpyAsync(len.data(), d_len.p, n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    mark("lengths");
    for (size_t k = 0; k < n; ++k) out_off[k + 1] = out_off[k] + len[k];
    const int64_t total = out_off[n];
    c->launches += 1;
    if (!out) return SCAR_OK;
    if (total > out_capacity) return scar_fail(SCAR_ERR_INVALID, "gather: %lld bytes needed", (long long)total);
    if (total == 0) return SCAR_OK;
    if ((rc = ensure(&c->d_gather_out, &c->gather_out_cap, (size_t)total))) return rc;
    d_out.p = c->d_gather_out;
    mark("output buffer");
    CU(cudaMemcpyAsync(d_off.p, out_off, n * 8, cudaMemcpyHostToDevice, s));
    gather_rows_kernel<<<grid, 256, 0, s>>>(c->d_text, c->d_fq_seq_off, c->d_fq_seq_len, (const int64_t *)d_rows.p,
                                            (const uint8_t *)d_flip.p, (const int64_t *)d_off.p, n_rows, c->platform, (uint8_t *)d_out.p);
    CU(cudaGetLastError());
    if (trace) { CU(cudaStreamSynchronize(s)); mark("gather kernel"); }
    CU(cudaMemcpyAsync(out, d_out.p, (size_t)total, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    mark("bytes to the host");
    if (trace) fprintf(stderr, "[scar gather] %lld rows, %lld bytes:%s\n", (long long)n_rows, (long long)total, timeline.c_str());
    c->launches += 1;
    return SCAR_OK;
}

extern "C" int scar_gunzip_host(const uint8_t *gz, int64_t gz_bytes, int32_t n_threads, int64_t chunk_bytes, uint8_t *out,
                                int64_t out_capacity, int64_t *out_bytes, int64_t *stats)
{
    if (!gz || gz_bytes < 0 || (!out && out_capacity > 0) || !out_bytes) return scar_fail(SCAR_ERR_INVALID, "bad argument");
    *out_bytes = 0;
    if (n_threads <= 0) n_threads = (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    scar_pgzip::ParallelGunzip pg(gz, gz_bytes, std::min(n_threads, 64), chunk_bytes);
    int64_t got = 0;
    uint8_t spill[4096];
    for (;;) {
        // (one more read once the buffer is full: the end of the stream, or text that does not fit)
        const bool full = got >= out_capacity;
        const int64_t r = full ? pg.read(spill, sizeof(spill)) : pg.read(out + got, std::min<int64_t>(out_capacity - got, 16 << 20));
        if (r < 0) return scar_fail(SCAR_ERR_INVALID, "%s", pg.error().c_str());
        if (r == 0) break;
        if (full) return scar_fail(SCAR_ERR_INVALID, "gzip: the text is longer than the output buffer (%lld bytes)", (long long)out_capacity);
        got += r;
    }
    *out_bytes = got;
    if (stats) { stats[0] = pg.chunks_total(); stats[1] = pg.chunks_guessed(); stats[2] = pg.chunks_redone(); stats[3] = pg.threads(); }
    return SCAR_OK;
}

extern "C" void scar_fastq_run_release(scar_fastq_run *run)
{
    if (!run) return;
    delete static_cast<RunOwner *>(run->owner);
    memset(run, 0, sizeof(*run));
}

extern "C" int scar_process_fastq_file(scar_ctx *c, const char *path, int64_t record_limit, uint64_t first_index,
                                       int32_t n_threads, int32_t keep_text, scar_fastq_run *run)
{
    if (!c || !path || !run) return scar_fail(SCAR_ERR_INVALID, "bad argument");
    memset(run, 0, sizeof(*run));
    const double t_begin = now_s();
    const bool trace = getenv("SCAR_INGEST_TRACE") != nullptr;
    auto mark = [&](const char *what) { if (trace) fprintf(stderr, "[scar ingest] %8.2f ms  %s\n", 1e3 * (now_s() - t_begin), what); };
    CU(cudaSetDevice(c->device));
    std::string err;
    MappedFile file;
    if (!file.open_ro(path, err)) return scar_fail(SCAR_ERR_INVALID, "FASTQ file: %s", err.c_str());
    RunOwner *own = new RunOwner();
    struct OwnerGuard { RunOwner *o; ~OwnerGuard() { delete o; } } og{own};
    const bool gz = file.n >= 2 && file.p[0] == 0x1f && file.p[1] == 0x8b;
    std::vector<BgzfBlock> blocks;
    const bool bgzf = gz && scan_bgzf(file.p, file.n, blocks);
    // (a single gzip stream is inflate-bound: up to 16 threads; copies and BGZF blocks stop gaining at 8)
    if (n_threads <= 0) n_threads = (int)std::min(gz && !bgzf ? 16u : 8u, std::max(1u, std::thread::hardware_concurrency()));
    n_threads = std::min(n_threads, 64);
    run->source_kind = bgzf ? 2 : gz ? 1 : 0;
    // a single gzip stream: inflated side by side (scar_pgzip.h) unless the file is small or SCAR_GZIP_SERIAL is set
    const bool gz_parallel = gz && !bgzf && n_threads > 1 && file.n >= (256 << 10) && !getenv("SCAR_GZIP_SERIAL");
    run->threads = gz && !bgzf && !gz_parallel ? 1 : n_threads;

    // the text size, to allocate once (every re-allocation in mid-stream is a cudaMalloc + copy + cudaFree with both
    // streams idle): the file itself; BGZF: the sum of the blocks' sizes; gzip: the length field of the last trailer
    // (exact for one member under 4 GiB; otherwise a typical FASTQ ratio) - grown on demand when the guess was low
    int64_t text_hint = file.n;
    if (bgzf) { text_hint = 0; for (const BgzfBlock &b : blocks) text_hint += b.out_len; }
    else if (gz) {
        const uint8_t *t = file.p + file.n - 4;
        const int64_t isize = file.n >= 18 ? (int64_t)((uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24)) : 0;
        text_hint = isize > file.n && isize < 1100 * file.n ? isize : 5 * file.n;
    }
    // ---- staging pieces: pinned, allocated when the producer first needs them (a small file pays for what it uses),
    // kept in the context between calls.  16 MB keeps the per-piece overheads (two stream synchronisations, ten
    // launches) under the copy time; a file smaller than that takes one piece of its own size.
    // Measured on a 3.46 GB text (10 M reads, profiles/r02/ingest_piece_sizes.txt): 16 / 32 / 64 / 128 MB pieces take
    // 0.76 / 0.62 / 0.65 / 0.71 s for the whole file - 32 MB once the text is large.
    int64_t piece = text_hint >= (512ll << 20) ? 32ll << 20 : 16ll << 20;
    if (const char *e = getenv("SCAR_FASTQ_FILE_PIECE_BYTES")) { const long long v = atoll(e); if (v > 0) piece = v; }
    else if (!gz) piece = std::min(piece, std::max<int64_t>(file.n, 1));
    else piece = std::min(piece, std::max<int64_t>(16 * file.n, 1 << 20));
    piece = std::max<int64_t>(1 << 17, (piece + 4095) / 4096 * 4096);       // (a BGZF block inflates to at most 64 KB)
    Stream S;
    S.cap = piece;
    if (c->h_piece_bytes < piece) {
        for (uint8_t *&p : c->h_piece) { if (p) cudaFreeHost(p); p = nullptr; }
        c->h_piece_bytes = piece;
    }
    S.device = c->device;
    S.pinned = c->h_piece;
    S.pinned_bytes = c->h_piece_bytes;
    mark("file mapped, pinned pieces ready");

    HostBuf &B = c->buf[0];
    int rc = scar_ensure_hostbuf(c, B, 1);
    if (rc) return rc;
    if ((rc = scar_ensure_hostbuf(c, c->buf[1], 1))) return rc;
    cudaStream_t s = B.stream, sc = c->buf[1].stream;
    cudaEvent_t copied = c->buf[1].done;
    c->fq_records_n = 0;
    if ((rc = ensure(&c->d_tile_counts, &c->tile_counts_cap, (size_t)((piece + 4095) / 4096)))) return rc;
    if ((rc = ensure(&c->d_tile_offsets, &c->tile_offsets_cap, (size_t)scar_fastq_offsets_elems(piece)))) return rc;
    if ((rc = ensure_keep(&c->d_text, &c->text_cap, (size_t)text_hint + 4096, 0))) return rc;

    mark("device buffers ready");
    HostText inflated;
    const bool keep = keep_text != 0;
    Workers W(std::max(0, n_threads - 1));
    std::thread producer([&] {
        if (file.n == 0) { Piece *pc = S.acquire_free(); if (pc) { pc->len = 0; pc->last = true; S.publish(); } }
        else if (bgzf) produce_bgzf(S, W, file, blocks, keep ? &inflated : nullptr);
        else if (gz_parallel) produce_gzip_parallel(S, W, file, keep ? &inflated : nullptr, n_threads);
        else if (gz) produce_gzip(S, W, file, keep ? &inflated : nullptr);
        else produce_plain(S, W, file);
    });
    struct Joiner { Stream &S; std::thread &t; ~Joiner() { S.request_stop(); if (t.joinable()) t.join(); } } joiner{S, producer};
    auto drain = [&](int code) { cudaStreamSynchronize(sc); cudaStreamSynchronize(s); return code; };

    const int want = c->platform == SCAR_PLATFORM_ILLUMINA;
    // the text stays resident: a cap below the device's own limit (and the way the tests reach the out-of-memory exit)
    int64_t text_limit = INT64_MAX;
    if (const char *e = getenv("SCAR_FASTQ_FILE_MAX_TEXT_BYTES")) { const long long v = atoll(e); if (v > 0) text_limit = v; }
    int64_t n_bytes = 0, nl_total = 0, rec_done = 0;
    uint8_t last_byte = '\n';
    int32_t *d_maxlen = reinterpret_cast<int32_t *>(c->d_acc + 4);           // spare accumulator word (after export_n)
    CU(cudaMemsetAsync(d_maxlen, 0, sizeof(unsigned long long), s));
    int32_t max_len = 0;
    double wait_s = 0.0;
    bool last = false;
    while (!last) {
        const double t0 = now_s();
        Piece *pc = S.wait_filled();
        wait_s += now_s() - t0;
        if (!S.error.empty()) return drain(scar_fail(SCAR_ERR_INVALID, "FASTQ file %s: %s", path, S.error.c_str()));
        last = pc->last;
        const int64_t len = pc->len, b0 = n_bytes;
        if (len > 0) {
            if (b0 + len > text_limit)
                return drain(scar_fail(SCAR_ERR_NOMEM, "FASTQ text larger than SCAR_FASTQ_FILE_MAX_TEXT_BYTES=%lld", (long long)text_limit));
            if ((size_t)(b0 + len) + 64 > c->text_cap) {                       // grow the device text (rare: streams idle first)
                cudaStreamSynchronize(sc); cudaStreamSynchronize(s);
                if ((rc = ensure_keep(&c->d_text, &c->text_cap, (size_t)((double)(b0 + len) * 1.5) + 4096, (size_t)b0))) return drain(rc);
            }
            CU(cudaMemcpyAsync(c->d_text + b0, pc->buf, (size_t)len, cudaMemcpyHostToDevice, sc));
            last_byte = pc->buf[len - 1];
            n_bytes += len;
        }
        if (last) CU(cudaMemsetAsync(c->d_text + n_bytes, 0, 64, sc));
        CU(cudaEventRecord(copied, sc));
        CU(cudaStreamWaitEvent(s, copied, 0));
        unsigned long long n_nl = 0;
        if (len > 0) {
            CU(scar_fastq_index_launch(c->d_text + b0, len, c->d_tile_counts, c->d_tile_offsets, c->d_export_n, s));
            c->launches += 3;
            CU(cudaMemcpyAsync(&n_nl, c->d_export_n, sizeof(n_nl), cudaMemcpyDeviceToHost, s));
        }
        CU(cudaStreamSynchronize(s));                                          // (the piece has been copied: its buffer is free)
        S.release();
        if (len > 0) {
            size_t want_nl = (size_t)(nl_total + (int64_t)n_nl) + 1;
            if (want_nl > c->nl_cap) {
                // lines of the whole text at the density seen so far (+ 3 %), else half as much again
                const double scale = n_bytes > 0 && text_hint > n_bytes ? 1.03 * (double)text_hint / (double)n_bytes : 1.5;
                want_nl = (size_t)((double)want_nl * (scale >= 1.0 ? scale : 1.5)) + 1024;
            }
            if ((rc = ensure_keep(&c->d_nl, &c->nl_cap, want_nl, (size_t)nl_total))) return drain(rc);
            CU(scar_fastq_positions_launch(c->d_text + b0, len, c->d_tile_offsets, c->d_nl + nl_total, (int64_t)n_nl, b0, s));
            ++c->launches;
            nl_total += (int64_t)n_nl;
        }
        // whole records only; the file's last record may lack its final newline
        int64_t rec_avail = nl_total / 4;
        if (last && nl_total % 4 == 3 && n_bytes > 0 && last_byte != '\n') ++rec_avail;
        if (record_limit > 0 && rec_avail >= record_limit) { rec_avail = record_limit; last = true; }
        const int64_t m = rec_avail - rec_done;
        if (m <= 0) continue;
        if (m > B.reads_cap && (rc = scar_ensure_hostbuf(c, B, m + m / 8 + 1024))) return drain(rc);
        size_t want_rec = (size_t)rec_avail;
        if (want_rec > c->fq_records_cap) {
            const double scale = n_bytes > 0 && text_hint > n_bytes ? 1.03 * (double)text_hint / (double)n_bytes : 1.5;
            want_rec = (size_t)((double)want_rec * (scale >= 1.0 ? scale : 1.5)) + 1024;
        }
        if ((rc = ensure_keep(&c->d_fq_records, &c->fq_records_cap, want_rec, (size_t)rec_done))) return drain(rc);
        if ((rc = ensure_keep(&c->d_fq_seq_off, &c->fq_seq_off_cap, want_rec, (size_t)rec_done))) return drain(rc);
        if ((rc = ensure_keep(&c->d_fq_seq_len, &c->fq_seq_len_cap, want_rec, (size_t)rec_done))) return drain(rc);
        CU(scar_fastq_records_launch(c->d_text, n_bytes, c->d_nl, nl_total, rec_done, m, want, c->d_fq_seq_off + rec_done,
                                     c->d_fq_seq_len + rec_done, B.f0_off, B.f0_len, B.f1_off, B.f1_len, c->d_status, c->sm_count, s,
                                     d_maxlen));
        ++c->launches;
        // the longest read so far and the number of scar patterns so far: re-plan / grow before this piece runs
        int32_t seen = 0;
        CU(cudaMemcpyAsync(&seen, d_maxlen, sizeof(seen), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        const int stored = c->platform == SCAR_PLATFORM_TRUSEQ ? std::max(0, seen - 12) : seen;
        max_len = std::max(max_len, seen);
        if (stored > c->max_read_len && (rc = scar_grow_read_len(c, stored))) return drain(rc);
        if ((rc = scar_table_reserve(c, m))) return drain(rc);
        scar_ragged dseq{c->d_text, c->d_fq_seq_off + rec_done, c->d_fq_seq_len + rec_done, 0};
        scar_ragged df0{c->d_text, B.f0_off, B.f0_len, 0}, df1{c->d_text, B.f1_off, B.f1_len, 0};
        if ((rc = scar_run_dev(c, &dseq, want ? &df0 : nullptr, want ? &df1 : nullptr, m, first_index + (uint64_t)rec_done,
                               B.cells, B.lens, B.samples, c->d_fq_records + rec_done, s))) return drain(rc);
        rec_done = rec_avail;
    }
    mark("last piece enqueued");
    S.request_stop();
    producer.join();
    CU(cudaStreamSynchronize(s));
    CU(cudaStreamSynchronize(sc));
    mark("streams drained");
    c->fq_records_n = rec_done;
    int32_t st = 0;
    CU(cudaMemcpy(&st, c->d_status, sizeof(st), cudaMemcpyDeviceToHost));
    if (st & 24) {
        CU(cudaMemset(c->d_status, 0, sizeof(int32_t)));
        if (st & 8) return scar_fail(SCAR_ERR_INVALID, "FASTQ: sequence and quality scores of different lengths");
        return scar_fail(SCAR_ERR_INVALID, "FASTQ: a header has no '+' between the index reads (the reference raises IndexError)");
    }
    if ((rc = scar_check_status(c))) return rc;
    if (c->verify && rec_done > 0) {
        // exact-or-fail scar-pattern identity: every scarred read's full key against the first read of its table entry
        scar_ragged dall{c->d_text, c->d_fq_seq_off, c->d_fq_seq_len, 0};
        int64_t ncol = 0;
        if ((rc = scar_verify_dev(c, &dall, c->d_fq_records, rec_done, first_index, &ncol, s))) return rc;
    }
    mark("verified");
    // ---- results to the host
    const size_t nrec = (size_t)std::max<int64_t>(rec_done, 1);
    own->records = (scar_record *)malloc(nrec * sizeof(scar_record));
    own->seq_off = (int64_t *)malloc(nrec * sizeof(int64_t));
    own->seq_len = (int32_t *)malloc(nrec * sizeof(int32_t));
    if (!own->records || !own->seq_off || !own->seq_len) return scar_fail(SCAR_ERR_NOMEM, "out of host memory for %lld records", (long long)rec_done);
    if (rec_done) {
        CU(cudaMemcpy(own->records, c->d_fq_records, (size_t)rec_done * sizeof(scar_record), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(own->seq_off, c->d_fq_seq_off, (size_t)rec_done * sizeof(int64_t), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(own->seq_len, c->d_fq_seq_len, (size_t)rec_done * sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    run->n_records = rec_done;
    run->text_bytes = n_bytes;
    run->records = own->records; run->seq_off = own->seq_off; run->seq_len = own->seq_len;
    if (keep) {
        if (gz) { own->text_heap = inflated.p; inflated.p = nullptr; run->text = own->text_heap; }
        else { own->text_map.p = file.p; own->text_map.n = file.n; own->text_map.fd = file.fd; file.p = nullptr; file.n = 0; file.fd = -1; run->text = own->text_map.p; }
    }
    run->max_seq_len = max_len;
    run->wait_seconds = wait_s;
    run->total_seconds = now_s() - t_begin;
    run->owner = own;
    og.o = nullptr;
    mark("results on the host");
    return SCAR_OK;
}
