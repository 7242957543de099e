/*
 * scar_b200.h — C ABI of libscar_b200.so: the B200 (sm_100a) implementation of ScarMapper's
 * per-read scar-classification path.  This is the drop-in boundary: everything the reference's
 * two hot call sites need, as plain pointers and sizes (no torch / C++ types).
 *
 * Reference interfaces replaced (paths under the reference tree):
 *   scarmapper/SlidingWindow.pyx:20-22      sliding_window(...)            -> scar_classify_dev / scar_process_host
 *   scarmapper/INDEL_Processing.py:135-196  ScarSearch.data_processing loop -> scar_process_host + scar_table_read
 *   scarmapper/INDEL_Processing.py:891-1024 consensus_demultiplex loop      -> scar_demux_dev (+ phasing in scar_classify_dev)
 *   scarmapper/INDEL_Processing.py:1126-1201 index_matching                 -> scar_demux_dev
 *   Valkyries/Sequence_Magic.py:123-136     match_maker(query, unknown)     -> scar_match_maker_batch
 *   scarmapper/INDEL_Processing.py:550-700  homeology_search / templated_insertion_search -> scar_homeology_batch
 *   Valkyries/FASTQ_Tools.py:562-653        FASTQ_Reader (text / gzip file)  -> scar_process_fastq_file
 *   scarmapper/INDEL_Processing.py:233-478  frequency_output (per-pattern body, rows) -> scar_frequency_rows_host
 *   scarmapper/INDEL_Processing.py:702-757  raw_data_output (body)         -> scar_raw_rows_host
 *   scarmapper/INDEL_Processing.py:57-80,759-795 window_mapping/cutsite_search -> scar_ctx_create (tables), scar_cutsite_search
 *   scarmapper/AlignmentProcessing.pyx:18-149 alignment_processing          -> scar_alignment_batch
 *   scarmapper/INDEL_Processing.py:876-889,988-1002 demultiplexed FASTQ      -> scar_demux_fastq_rows_host
 *
 * Conventions: every function returns 0 on success or a negative SCAR_ERR_* code;
 * scar_last_error() gives the text for the calling thread.  No exceptions cross the boundary.
 * The caller owns every buffer.  "_dev" functions take DEVICE pointers and a cudaStream_t
 * (passed as void*; NULL = default stream) and are asynchronous on that stream; "_host"
 * functions take HOST pointers and do their own host<->device copies.  A context is bound to one
 * device and is thread-compatible (one thread at a time).  There is no CPU fallback: every entry
 * point that computes fails with SCAR_ERR_CUDA when no device is usable.
 */
#ifndef SCAR_B200_H
#define SCAR_B200_H

#include <stdint.h>
#include "scar_records.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
    SCAR_OK = 0,
    SCAR_ERR_CUDA = -1,          /* CUDA runtime error (no device, launch failure, ...) */
    SCAR_ERR_INVALID = -2,       /* bad argument */
    SCAR_ERR_UNSUPPORTED = -3,   /* input outside the supported envelope */
    SCAR_ERR_TABLE_FULL = -4,    /* scar table capacity exhausted (grow and retry) */
    SCAR_ERR_NOMEM = -5,
    SCAR_ERR_HASH_COLLISION = -6 /* verification found two keys under one 128-bit tag */
};

enum { SCAR_PLATFORM_ILLUMINA = 0, SCAR_PLATFORM_TRUSEQ = 1, SCAR_PLATFORM_RAMSDEN = 2 };

/* A list of byte strings, three layouts:
 *   CSR     offsets != NULL, lengths == NULL: item i = bytes[offsets[i] .. offsets[i+1])  (n+1 offsets)
 *   views   offsets != NULL, lengths != NULL: item i = bytes[offsets[i] .. offsets[i]+lengths[i]), offsets
 *           non-decreasing (fields inside a FASTQ text buffer, see scar_fastq_index)
 *   stride  offsets == NULL: item i starts at i*stride and is lengths[i] long, or `stride` long when
 *           lengths == NULL. */
/* DEVICE buffers handed to the _dev entry points must be readable for 16 bytes past the last item (the kernels
 * use aligned 4/8-byte loads that may straddle the end of an item). */
typedef struct scar_ragged {
    const uint8_t *bytes;
    const int64_t *offsets;
    const int32_t *lengths;
    int64_t stride;
} scar_ragged;

/* Static description of a run: what ScarSearch.__init__/window_mapping/cutsite_search,
 * DataProcessing.dictionary_build and TargetMapper.phasing compute once per run. */
typedef struct scar_config {
    int32_t device;               /* CUDA device ordinal */
    int32_t platform;             /* SCAR_PLATFORM_* (args.Platform) */
    int32_t mismatch;             /* index match threshold; -1 = platform default 1/0/3 (INDEL_Processing.py:1142-1147) */
    int32_t n_targets;
    const uint8_t *target_bytes;  /* upper-cased target regions, concatenated */
    const int64_t *target_off;    /* [n_targets+1] */
    const int32_t *target_cutsite;/* [n_targets] (scar_cutsite_search) */
    int32_t n_samples;            /* manifest rows, manifest order */
    const int32_t *sample_target; /* [n_samples] target id of each sample (index_dict[s][7]) */
    const uint8_t *left_index_bytes;  /* index_dict[s][0]: Master_Index_File column 3, upper-cased */
    const int64_t *left_index_off;    /* [n_samples+1] */
    const uint8_t *right_index_bytes; /* index_dict[s][2]: Master_Index_File column 2, upper-cased */
    const int64_t *right_index_off;   /* [n_samples+1] */
    /* TargetMapper.phasing lists, concatenated over targets; target t owns phases
     * [phase_off[t], phase_off[t+1]).  NULL phase_off = no phasing. */
    const int32_t *phase_off;         /* [n_targets+1] */
    const uint8_t *phase_r1_bytes; const int64_t *phase_r1_off;
    const uint8_t *phase_r2_bytes; const int64_t *phase_r2_off;
    const uint8_t *hr_donor; int32_t hr_len;   /* args.HR_Donor ("" = off) */
    double n_limit;               /* args.N_Limit */
    int32_t min_length;           /* args.Minimum_Length */
    int32_t max_read_len;         /* longest read the context must accept (<= 1024) */
    int64_t table_capacity;       /* scar-table slots, power of two; 0 = default */
    /* per-read drop-in only: a 10-nt cutwindow per target for the SlidingWindow.pyx:56-60 test
     * (the driver always passes 8 nt, which can never match; NULL = none) */
    const uint8_t *cutwindow_bytes;   /* [n_targets][10] or NULL */
} scar_config;

typedef struct scar_ctx scar_ctx;

const char *scar_last_error(void);
int scar_version(void);

/* ScarSearch.cutsite_search (INDEL_Processing.py:759-795) on the host; returns the cutsite in
 * *cutsite, or SCAR_ERR_INVALID when the sgRNA does not map (the reference exits). */
int scar_cutsite_search(const uint8_t *target, int32_t len, const uint8_t *sgrna, int32_t sg_len,
                        int32_t reverse_comp, int32_t *cutsite);

int scar_ctx_create(const scar_config *cfg, scar_ctx **out);
void scar_ctx_destroy(scar_ctx *ctx);

/* Layout of the packed read batch in HBM: W = scar_cells_per_read() 64-bit cells per read, stored
 * CELL-MAJOR: cell j of read r of an n-read batch is cells[j * n + r] (2-bit codes | ACGT mask << 32 |
 * N mask << 48 for bases 16j .. 16j+15). */
int scar_cells_per_read(const scar_ctx *ctx);
int scar_window_counts(const scar_ctx *ctx, int32_t target, int32_t *n_left, int32_t *n_right);
/* size of the static blob K3 stages per CTA, whether it fits shared memory, and how many loci use the
 * anchor-and-extend search (unique 10-mers) */
int scar_ctx_info(const scar_ctx *ctx, int64_t *static_bytes, int32_t *tables_in_smem, int32_t *n_anchor_targets);

/* how K3 is launched for this context: whether a tile's cells are staged in shared memory, resident CTAs per SM and
 * dynamic shared memory per CTA (bench.py reports them per config) */
int scar_k3_plan(const scar_ctx *ctx, int32_t *cells_staged, int32_t *ctas_per_sm, int32_t *smem_bytes);

/* K1: raw bytes -> packed rows (2-bit codes + ACGT mask + N mask), applying the platform's stored-
 * sequence transform (Illumina as is, TruSeq seq[6:-6], Ramsden rcomp; INDEL_Processing.py:946,978,981). */
int scar_pack_dev(scar_ctx *ctx, const scar_ragged *seq, int64_t n_reads,
                  uint64_t *cells, uint32_t *lens, void *stream);   /* lens[r] = stored length | N count << 16 */

/* K2: sample assignment per read (manifest position or SCAR_SAMPLE_NONE).
 * Illumina: f0/f1 = name.split(":")[-1].split("+")[0..1]; TruSeq/Ramsden: taken from seq. */
int scar_demux_dev(scar_ctx *ctx, const scar_ragged *seq, const scar_ragged *f0, const scar_ragged *f1,
                   int64_t n_reads, uint16_t *samples, void *stream);

/* K3: phasing + filters + sliding-window junction search + junction call. */
int scar_classify_dev(scar_ctx *ctx, const uint64_t *cells, const uint32_t *lens,
                      const uint16_t *samples, int64_t n_reads, scar_record *records, void *stream);

/* K4: accumulate per-sample counters, phase histograms and the scar table.  first_index is the
 * global index of read 0 of this batch (first-seen order across batches and ranks). */
int scar_aggregate_dev(scar_ctx *ctx, const scar_ragged *seq, const scar_record *records,
                       int64_t n_reads, uint64_t first_index, void *stream);

/* K1 + K3 (+ K4 when aggregate != 0) FUSED: one pass over the RAW read bytes.  A CTA stages a tile's raw bytes in
 * shared memory (TMA bulk copies), packs them there, classifies, writes the records and - with `aggregate` - updates
 * the per-sample counters, phase histograms and the scar table, so the packed batch never travels through HBM and
 * the raw reads are read once.  Same records / tables as scar_pack_dev + scar_classify_dev + scar_aggregate_dev.
 * Takes reads up to 320 nt on the Illumina and TruSeq platforms (scar_fused_info: eligible); SCAR_ERR_UNSUPPORTED
 * otherwise. */
int scar_classify_raw_dev(scar_ctx *ctx, const scar_ragged *seq, const uint16_t *samples, int64_t n_reads,
                          uint64_t first_index, scar_record *records, int32_t aggregate, void *stream);
int scar_fused_info(const scar_ctx *ctx, int32_t *eligible, int32_t *ctas_per_sm, int32_t *smem_bytes);

/* Whole path over a batch already resident in HBM on `stream`: K2, then the fused kernel (scar_classify_raw_dev with
 * aggregate = 1) when the context is eligible for it, else K1 -> K3 -> K4.  cells/lens/samples/records are
 * caller-provided device scratch/outputs (cells and lens are untouched on the fused path; SCAR_NO_FUSE=1 forces
 * the three-kernel path). */
int scar_run_dev(scar_ctx *ctx, const scar_ragged *seq, const scar_ragged *f0, const scar_ragged *f1,
                 int64_t n_reads, uint64_t first_index, uint64_t *cells, uint32_t *lens,
                 uint16_t *samples, scar_record *records, void *stream);

/* Whole path from HOST buffers: chunks the batch, overlaps H2D / kernels / D2H on internal
 * streams.  records may be NULL (tables and counters only). */
int scar_process_host(scar_ctx *ctx, const scar_ragged *seq, const scar_ragged *f0, const scar_ragged *f1,
                      int64_t n_reads, uint64_t first_index, scar_record *records);

/* Results accumulated so far (synchronises the context's streams). */
int scar_counters_read(scar_ctx *ctx, int64_t *counters /* [n_samples][SCAR_N_COUNTERS] */,
                       int64_t *unassigned);
int scar_phase_hist_read(scar_ctx *ctx, int64_t *hist /* [n_samples][2][SCAR_PHASE_BINS] */);
#define SCAR_PHASE_BINS 34   /* bin 0 = "No Read N Phasing", bin 1+k = phase k, last = n/a */
int scar_table_size(scar_ctx *ctx, int64_t *n_entries);
/* entries sorted by (sample, first_idx): the reference's first-seen dict order per sample. */
int scar_table_read(scar_ctx *ctx, scar_entry *entries, int64_t capacity, int64_t *n_entries);
int scar_verify_dev(scar_ctx *ctx, const scar_ragged *seq, const scar_record *records, int64_t n_reads,
                    uint64_t first_index, int64_t *n_collisions, void *stream);
int scar_reset(scar_ctx *ctx);
int scar_reset_dev(scar_ctx *ctx, void *stream);      /* asynchronous on `stream` */
/* Exact-or-fail key identity for scar_process_host (off by default).  A scar-table entry is identified by a 128-bit
 * tag of its key; with verification on, scar_process_host keeps the whole batch resident as ONE chunk and afterwards
 * compares every scarred read's full key (sample, junctions, kind, segment bytes) with the first read of its entry
 * (scar_verify_dev): a mismatch fails the call with SCAR_ERR_HASH_COLLISION instead of merging two keys silently.
 * Needs ~ (raw bytes + 48) per read of device memory for the batch (scar_device_memory). */
int scar_ctx_set_verify(scar_ctx *ctx, int32_t on);
int scar_device_memory(int32_t device, int64_t *free_bytes, int64_t *total_bytes);

/* Size the persistent grids of this context's kernels for n_sms SMs instead of all of them (0 = all): a multi-GPU
 * job leaves a few SMs to the table merge and the NCCL kernels that run beside the next batch. */
int scar_ctx_set_sm_limit(scar_ctx *ctx, int32_t n_sms);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t scar_launch_count(const scar_ctx *ctx);

/* Multi-GPU: export this rank's table, merge another rank's into it (count SUM, first_idx MIN). */
int scar_table_export_dev(scar_ctx *ctx, scar_entry *entries, int64_t capacity, int64_t *n_entries, void *stream);
int scar_table_merge_dev(scar_ctx *ctx, const scar_entry *entries, int64_t n_entries, void *stream);
/* Owner-partitioned merge (cost independent of the GPU count): region o of `regions` ([n_parts][region_capacity]
 * entries; entry 0 of a region is a header whose h1 is the count) receives this rank's entries owned by rank o
 * (owner = (h2 >> 40) % n_parts).  Exchange the regions with one all-to-all, scar_table_clear_dev, then fold the
 * received regions: every rank ends up holding the merged entries it owns.  All three are stream-ordered. */
int scar_table_export_parts_dev(scar_ctx *ctx, scar_entry *regions, int32_t n_parts, int64_t region_capacity, void *stream);
/* The export fused with the exchange over NVLink peer memory: peer_regions[o] (host array of n_parts DEVICE
 * pointers) is rank o's receive buffer ([n_parts][region_capacity] entries) mapped into this process (CUDA IPC /
 * symmetric memory; peer_regions[my_rank] is this rank's own).  Every entry is stored directly into region
 * `my_rank` of its owner's buffer and the count is published as that region's header - no all-to-all pass.
 * The caller orders the ranks (a barrier before: the peers have consumed their buffers; a barrier after: all
 * stores have landed), then scar_table_clear_dev + scar_table_merge_parts_dev on its own receive buffer. */
int scar_table_export_p2p_dev(scar_ctx *ctx, void *const *peer_regions, int32_t n_parts, int32_t my_rank,
                              int64_t region_capacity, void *stream);
int scar_table_clear_dev(scar_ctx *ctx, void *stream);
int scar_table_merge_parts_dev(scar_ctx *ctx, const scar_entry *regions, int32_t n_parts, int64_t region_capacity, void *stream);

/* Dense accumulators as one int64 block [unassigned, counters[n_samples][16], phase_hist[n_samples][2][34]]
 * for an all-reduce(SUM) across ranks: export to / import from caller-owned device memory. */
int64_t scar_accumulator_words(const scar_ctx *ctx);
int scar_accumulators_export_dev(scar_ctx *ctx, int64_t *out, void *stream);
int scar_accumulators_import_dev(scar_ctx *ctx, const int64_t *in, void *stream);

/* FASTQ_Reader.seq_read + the header split of index_matching over a text buffer (host only): emits
 * (offset, length) views into `text` for the sequence and, when want_index_fields, for
 * name.split(":")[-1].split("+")[0] and [1].  Parses whole records only; *consumed is where the next
 * call must resume (is_final accepts a last record without a newline).  Returns SCAR_ERR_INVALID where
 * the reference raises (sequence / quality length mismatch, header without '+'). */
int scar_fastq_index(const uint8_t *text, int64_t n_bytes, int32_t is_final, int32_t want_index_fields,
                     int64_t capacity, int64_t *seq_off, int32_t *seq_len, int64_t *f0_off, int32_t *f0_len,
                     int64_t *f1_off, int32_t *f1_len, int64_t *n_records, int64_t *consumed);

/* The same from FASTQ TEXT in host memory (whole records; a last record may lack its newline): the text is
 * copied to the GPU once, in pieces (256 MB; SCAR_FASTQ_PIECE_BYTES overrides) on one stream, while a second
 * stream indexes each piece there (newline count / scan / positions, one thread per record for the views)
 * and runs the records that have become complete through the whole path, on the views.  records (may be
 * NULL) receives one record per FASTQ record in file order; *n_records the count. */
int scar_process_fastq_host(scar_ctx *ctx, const uint8_t *text, int64_t n_bytes, uint64_t first_index,
                            scar_record *records, int64_t records_capacity, int64_t *n_records);

/* After scar_process_fastq_host(records = NULL): copy the records of that call to the host once their number
 * is known. */
int scar_fastq_records_fetch(scar_ctx *ctx, scar_record *records, int64_t n_records);

/* A whole FASTQ FILE through the path (replaces FASTQ_Reader as the loader in front of consensus_demultiplex,
 * Valkyries/FASTQ_Tools.py:562-653 + scarmapper/INDEL_Processing.py:891-1024): plain text, gzip (one stream or several
 * members) or BGZF, told apart by the magic bytes as the reference tells them apart by MIME type.  A producer thread
 * reads / inflates the file into pinned pieces (BGZF blocks and plain text by n_threads threads side by side; a
 * single DEFLATE stream by one thread, overlapped with everything else) while the calling thread copies each piece to
 * the GPU, indexes it there and runs the records that have become complete through the whole path.  The context
 * adapts to the run: longer reads than max_read_len re-plan it, the scar table grows (scar_table_reserve).
 * record_limit > 0 stops after that many records (--Verbose DEBUG).  With scar_ctx_set_verify the run is checked key
 * by key at the end.  keep_text != 0 keeps the (inflated) text for the caller.  The arrays of `run` belong to the
 * library until scar_fastq_run_release. */
typedef struct scar_fastq_run {
    int64_t n_records;             /* FASTQ records processed, file order */
    int64_t text_bytes;            /* (inflated) text consumed */
    const uint8_t *text;           /* the text (keep_text) or NULL */
    const int64_t *seq_off;        /* [n_records] views of the sequence lines into the text */
    const int32_t *seq_len;
    const scar_record *records;    /* [n_records] */
    int32_t max_seq_len;           /* longest sequence line */
    int32_t source_kind;           /* 0 plain text, 1 gzip stream, 2 BGZF */
    int32_t threads;               /* threads that read / inflated */
    int32_t reserved;
    double wait_seconds;           /* time the calling thread waited for the producer (inflate-bound when ~ total) */
    double total_seconds;
    void *owner;
} scar_fastq_run;
int scar_process_fastq_file(scar_ctx *ctx, const char *path, int64_t record_limit, uint64_t first_index,
                            int32_t n_threads, int32_t keep_text, scar_fastq_run *run);
void scar_fastq_run_release(scar_fastq_run *run);

/* Stored sequences (as scar_stored_gather_host: platform transform, row k reverse-complemented when flip[k] != 0) of
 * the reads `rows` of the run whose text is still resident on the device - the last scar_process_fastq_file of this
 * context - gathered THERE: only the chosen reads come back (the first read of every scar pattern for
 * frequency_output, scarmapper/INDEL_Processing.py:299-328), so a caller that needs no other text can run with
 * keep_text = 0.  out_off receives n_rows + 1 offsets; out = NULL sizes only. */
int scar_fastq_gather_host(scar_ctx *ctx, const int64_t *rows, int64_t n_rows, const uint8_t *flip, uint8_t *out,
                           int64_t out_capacity, int64_t *out_off);

/* The inflate in front of that loader, on its own: ONE gzip stream (or several members back to back, zero padding
 * allowed) -> text, by n_threads threads (<= 0: up to 8).  Replaces gzip.open() in FASTQ_Reader
 * (Valkyries/FASTQ_Tools.py:571-575).  The stream is cut into chunks of chunk_bytes compressed bytes (<= 0: chosen);
 * every chunk guesses a block start and decodes with an unknown 32 KiB window, a second pass resolves the windows in
 * order (csrc/scar_pgzip.h).  Exact or fail: every chunk must start on the bit where the one before it ended, every
 * member's CRC-32 and length are checked.  Host memory only, no CUDA.  stats (optional, 4 values): chunks, chunks whose
 * guessed start was taken, chunks decoded again, threads.  The text must fit out_capacity. */
int scar_gunzip_host(const uint8_t *gz, int64_t gz_bytes, int32_t n_threads, int64_t chunk_bytes, uint8_t *out,
                     int64_t out_capacity, int64_t *out_bytes, int64_t *stats);

/* Room for n_more further scar patterns: rebuilds the table at a larger power of two when it would pass half full
 * (entries re-inserted through the claimed-slot list; first-seen indices and counts kept).  Synchronises the device. */
int scar_table_reserve(scar_ctx *ctx, int64_t n_more);

/* The body of ScarSearch.raw_data_output (scarmapper/INDEL_Processing.py:702-757), host side: one text line per
 * scarred read of `sample` (records[r].status == SCAR_ST_SCAR), in read order, built from the records and the raw
 * reads (host buffers; `platform` selects the stored-sequence transform), with the swaps and reverse complements
 * of 3'-strand guides (reverse_comp != 0).  Call with out = NULL to get the size, then with a buffer. */
int scar_raw_rows_host(const scar_record *records, const scar_ragged *reads, int64_t n, int32_t sample,
                       int32_t platform, const uint8_t *target_region, int32_t target_len, int32_t cutsite,
                       int32_t reverse_comp, uint8_t *out, int64_t out_capacity, int64_t *out_bytes);

/* The same for a LIST of reads: rows[k] is a read index (the scarred reads of one sample, in read order).  A run with
 * many samples buckets its reads once and pays O(reads of the sample) per Raw_Data file. */
int scar_raw_rows_indexed_host(const scar_record *records, const scar_ragged *reads, const int64_t *rows, int64_t n_rows,
                               int32_t platform, const uint8_t *target_region, int32_t target_len, int32_t cutsite,
                               int32_t reverse_comp, uint8_t *out, int64_t out_capacity, int64_t *out_bytes);

/* Stored sequences (what ScarSearch is handed per platform, scarmapper/INDEL_Processing.py:946,978,981) of a list of
 * reads, gathered into one CSR buffer; row k is reverse-complemented when flip[k] != 0 (frequency_output works on
 * rcomp(consensus) for 3'-strand guides, :312-327).  out_off receives n_rows + 1 offsets; out = NULL sizes only. */
int scar_stored_gather_host(const scar_ragged *reads, const int64_t *rows, int64_t n_rows, int32_t platform,
                            const uint8_t *flip, uint8_t *out, int64_t out_capacity, int64_t *out_off);

/* The demultiplexed FASTQ of --Demultiplex True (scarmapper/INDEL_Processing.py:876-889,988-1002;
 * Valkyries/FASTQ_Tools.py:532-551): "@name\nseq\n+\nqual\n" for the reads rows[0..n_rows) (one sample's reads in file
 * order), formatted straight from the FASTQ text; seq_off / seq_len are the sequence views scar_fastq_index reports.
 * name = header.strip("\n").strip("@"), seq and qual stripped of white space, as FASTQ_Reader.seq_read delivers them.
 * out = NULL sizes only. */
int scar_demux_fastq_rows_host(const uint8_t *text, int64_t n_bytes, const int64_t *seq_off, const int32_t *seq_len,
                               const int64_t *rows, int64_t n_rows, uint8_t *out, int64_t out_capacity, int64_t *out_bytes);

/* ScarSearch.homeology_search (scarmapper/INDEL_Processing.py:550-657) and templated_insertion_search (:659-700)
 * for a batch of scar patterns on the device (host buffers), one thread per key.  Key i: consensus_i, its
 * junctions clj_i / crj_i, insertion_i, the target junctions tlj_i / trj_i as frequency_output passes them (i.e.
 * already swapped for 3'-strand guides) and target_idx_i into `targets` (the sequence homeology_search is given:
 * the target region, reverse-complemented for 3'-strand guides) and `regions` (self.target_region as stored,
 * which templated_insertion_search slices).  out[i] = 16 int32 (struct ScarHomeology in
 * scarmapper_b200/csrc/scar_homeology_core.h): left and right homeology {distance 0/1 or -1, position, query
 * start/length in the consensus, segment start/length in the target} and the start of the left / right 5-nt
 * template in the region (-1 = none; searched when the insertion has at least 5 nt). */
int scar_homeology_batch(int32_t device, const scar_ragged *consensus, const scar_ragged *insertion,
                         const int32_t *clj, const int32_t *crj, const int32_t *tlj, const int32_t *trj,
                         const int32_t *target_idx, const scar_ragged *targets, const scar_ragged *regions,
                         int32_t n_targets, int64_t n, int32_t *out);

/* AlignmentProcessing.alignment_processing (scarmapper/AlignmentProcessing.pyx:18-149; legacy module of the reference:
 * walks a gapped alignment of the consensus against the genome) for a batch of alignments on the device (host
 * buffers), one thread per alignment.  Alignment i: gapped_genomic_i, gapped_consensus_i (one length, equal to
 * gap_con_len_i), cutposition_i, gap_genome_len_i (the 10-90 % window is taken from it, as in the reference).
 * counts[i] = 8 int32 (struct ScarAlignmentCounts, scarmapper_b200/csrc/scar_alignment_core.h): left / right deleted
 * columns, left / right inserted columns, number of microhomology / insertion / snv strings, overflow flag;
 * spans[i] = [3][max_spans][2] int32 (start, length) of those strings: microhomology slices of gapped_genomic,
 * insertion and snv slices of gapped_consensus. */
int scar_alignment_batch(int32_t device, const scar_ragged *genomic, const scar_ragged *consensus,
                         const int32_t *cutposition, const int32_t *gap_con_len, const int32_t *gap_genome_len,
                         int64_t n, int32_t max_spans, int32_t *counts, int32_t *spans);

/* The per-pattern body of ScarSearch.frequency_output (scarmapper/INDEL_Processing.py:233-478) for ONE sample, host side:
 * scar-type classification (HR / SNV / TMEJ / NHEJ / Non-MH Deletion / Insertion / Other), junction-type totals, SNV
 * tallies, PatternThreshold filter, ordering (count, then first-seen rank, both descending) and the text rows of the
 * *_ScarMapper_Frequency.txt file (26 tab-separated columns; the Frequency column is Python's str(count / total)).
 * Inputs: what the GPU pass leaves per sample (scar_table_read order = first-seen) plus scar_homeology_batch's results. */
typedef struct scar_frequency_in {
    int64_t n;                     /* scar patterns of the sample, first-seen order */
    const uint32_t *counts;        /* [n] */
    const scar_record *first;      /* [n] record of the first read of each pattern */
    const uint8_t *reads;          /* stored sequences of those reads (scar_stored_gather_host), CSR */
    const int64_t *read_off;       /* [n + 1] */
    const int32_t *pattern;        /* [n][16] scar_homeology_batch output */
    const uint8_t *region; int32_t region_len;     /* ScarSearch.target_region */
    int32_t cutsite, reverse_comp; /* ScarSearch.cutsite; target_dict[..][5] == "YES" */
    int64_t passing, no_cut;       /* summary_data[1], summary_data[6][1] */
    double pattern_threshold;      /* args.PatternThreshold */
} scar_frequency_in;
/* junction_type_data[6] (TMEJ, NHEJ, Insertion, Other, Non-MH Deletion, SNV read totals); label_sums[7] /
 * label_order[7]: sum of count / total per scar type (HR, SNV, TMEJ, NHEJ, Non-MH Deletion, Insertion, Other) and the
 * types in order of first appearance in the sorted list (-1 padded); snv[4][10][4]: SNV tallies (left, left x count,
 * right, right x count) per k-mer position per nucleotide G, A, T, C; plot[n_plot][6]: per written row (type, pattern
 * frequency count / (passing - no_cut), left deletion, right deletion, microhomology size, insertion size);
 * totals[2]: number of patterns, number of scarred reads.  The rows are kept for the calling thread until
 * scar_frequency_rows_fetch copies them out (*rows_bytes says how many bytes). */
int scar_frequency_rows_host(const scar_frequency_in *in, int64_t *rows_bytes, int64_t *junction_type_data, double *label_sums,
                             int32_t *label_order, int64_t *snv, double *plot, int64_t plot_capacity, int64_t *n_plot,
                             int64_t *totals);
int scar_frequency_rows_fetch(uint8_t *rows, int64_t capacity);

/* Sequence_Magic.match_maker over a batch on the device (host buffers): out[i] = lev(q_i,u_i) - (|u_i|-|q_i|). */
int scar_match_maker_batch(int32_t device, const scar_ragged *query, const scar_ragged *unknown,
                           int64_t n, int32_t *out);

#ifdef __cplusplus
}
#endif
#endif
