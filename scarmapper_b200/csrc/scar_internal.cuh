// scar_internal.cuh — shared device/host definitions of libscar_b200 (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/scar_b200.h"
#include "scar_demux_core.h"
#include "scar_swar.h"

#define SCAR_KMER 10                 // hard-coded window (INDEL_Processing.py:67,76; SlidingWindow.pyx:49-50)
#define SCAR_KEY_MASK 0xFFFFFu       // 10 bases x 2 bits
#define SCAR_TABLE_EMPTY 0xFFFFFFFFu // table entry = key << 12 | i ; i = 4095 is reserved for the empty pattern
#define SCAR_MAX_WINDOWS 4095        // window index i is stored in 12 bits (0..4094)
#define SCAR_MAX_PHASES 32
#define SCAR_MAX_INDEX_LEN 64        // Myers bit-vector width
#define SCAR_MAX_CLASSES 16          // distinct byte values over all index sequences (+ class 0 = other)

// ---- base encoding --------------------------------------------------------------------------------
// code = (byte >> 1) & 3 : A=0 C=1 T=2 G=3 ; complement = code ^ 2.
__host__ __device__ __forceinline__ uint32_t scar_base_code(uint8_t b) { return (b >> 1) & 3u; }
__host__ __device__ __forceinline__ bool scar_base_valid(uint8_t b)
{
    return b == 'A' || b == 'C' || b == 'G' || b == 'T';
}

// A packed read is W 64-bit cells, stored CELL-MAJOR in HBM (cell j of read r at cells[j*stride + r],
// stride = reads in the batch) so that one-thread-per-read kernels load coalesced.  Cell j covers
// bases [16j, 16j+16):
//   bits  0..31  2-bit codes, base k of the cell at bits 2k..2k+1
//   bits 32..47  ACGT mask (1 = byte is one of A,C,G,T)
//   bits 48..63  N mask    (1 = byte is 'N')
// bases past the read's length have all three fields 0.

// A "10-mer -> payload" table in the static blob, entries key << 12 | payload (payload <= 4094, empty = 0xFFFFFFFF).
// Two layouts, one per context (K3Params.compact_tables):
//   bucketed  perfect hash with two entries per 8-byte bucket, multiplier searched so that no bucket overflows:
//             ONE probe (bucket = mulhi(t * mult, n), t = key << 12) - the fastest lookup, ~13 % full;
//   compact   hash-and-displace (h = t * mult; d = disp[mulhi(h, nb)]; slot = mulhi((h << 8) + d * unit, n)):
//             two dependent loads, ~72 % full - about a fifth of the shared memory, which is what lets several
//             groups of a CTA work beside the tables of a dozen loci.
struct ScarTableDesc {
    uint32_t off;       // byte offset in the blob of the buckets (uint2) / of the entries (u32)
    uint32_t mult;
    uint32_t n;         // buckets / slots
    uint32_t doff;      // compact: byte offset of the u16 displacements
    uint32_t nb;        // compact: displacement buckets
    uint32_t unit;      // compact: ceil(2^32 / n)
};

// the slot arithmetic of a compact table, shared by the host builder and the kernel
__host__ __device__ __forceinline__ uint32_t scar_mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
__host__ __device__ __forceinline__ uint32_t scar_chd_bucket(uint32_t h, uint32_t nb) { return scar_mulhi32(h, nb); }
__host__ __device__ __forceinline__ uint32_t scar_chd_slot(uint32_t h, uint32_t d, uint32_t unit, uint32_t n)
{
    return scar_mulhi32((h << 8) + d * unit, n);
}

struct ScarTargetMeta {
    int32_t cutsite, length;
    int32_t n_left, n_right;
    ScarTableDesc left, right;         // exhaustive scan: "10-mer -> smallest window index i" per side
    uint32_t cutwindow_key;            // SlidingWindow.pyx:56-60 test; 0xFFFFFFFF = none
    int32_t phase_off, n_phase;
    // anchor-and-extend path (locus is all ACGT and every 10-mer of it is unique): ONE table
    // "10-mer -> position t in the locus" and the locus itself 2-bit packed, 16 bases per word
    int32_t fast;                      // 1: use pos / tw_off, the left/right tables are not built
    ScarTableDesc pos;
    uint32_t tw_off;                   // byte offset in the blob of the packed locus (word 0 = 16 bases of padding)
    // phasing lookup (all non-empty phase strings of a side share one length <= 6): 2 x 4^n bytes in the
    // blob, entry = 1 + first matching k (0 = none), indexed by the prefix / suffix 2-bit code
    uint32_t ph_lut_off;               // 0 = no table, use the phase list
    int32_t ph_n1, ph_n2;
    int32_t pad;
};

struct ScarPhase {
    uint32_t r1_code, r2_code;         // r2_code = codes of rcomp(R2 string): compared with the read's suffix
    uint8_t r1_len, r2_len;            // 0 = empty string; 0xFF = contains non-ACGT bytes (can never match)
    uint8_t pad[2];
};

// 64-bit mixers (splitmix64 / murmur3 finalisers) for the scar-key identity
__host__ __device__ __forceinline__ uint64_t scar_mix1(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t scar_mix2(uint64_t k)
{
    k ^= k >> 33; k *= 0xFF51AFD7ED558CCDull;
    k ^= k >> 33; k *= 0xC4CEB9FE1A85EC53ull;
    return k ^ (k >> 33);
}

struct __align__(16) ScarTableSlot {    // 32 bytes, zero-initialised
    unsigned __int128 tag;              // h1 | h2 << 64 ; 0 = empty.  Claimed with ONE 128-bit CAS (ATOMG.CAS.128)
    unsigned long long inv_first;       // ~first_idx, combined with atomicMax
    uint32_t count;
    uint32_t sample;
};


// ---- kernel parameter blocks and launchers ---------------------------------------------------------
struct K1Params {
    scar_ragged seq;      // device pointers
    uint64_t *cells;      // cell-major [W][stride]
    int64_t stride;
    uint32_t *lens;       // [n] stored length | N count << 16
    int64_t n_reads;
    int32_t W;
    int32_t platform;
    int32_t *status;      // device flag: bit 0 set when a read is longer than 16*W
};

struct K2Params {
    scar_ragged seq, f0, f1;   // device pointers
    uint16_t *samples;         // [n] out
    int64_t n_reads;
    const uint8_t *class_map;  // [256] byte -> class (0 = appears in no index string)
    const DemuxString *right;  // distinct right_index strings (scored against u0)
    const DemuxString *left;   // distinct left_index strings (scored against u1)
    const uint64_t *peq;
    const uint16_t *first_sample; // [n_right][n_left] first manifest row with that pair, 0xFFFF = none
    const uint8_t *fast_codes;    // [(n_right + n_left)][16] 2-bit base codes of the distinct strings (fast only)
    int32_t n_right, n_left, n_classes;
    int32_t platform, mismatch;
    int32_t fast;              // every index string is ACGT-only and <= 16 nt: 2-bit Hamming shortcut
    int32_t pad;
    // direct tables (fast only, every distinct string of the side n <= 8 nt long): accepted-string mask of
    // every possible all-ACGT field of n bases, indexed by its 2-bit code; null when not applicable
    const uint64_t *lut_right, *lut_left;
    int32_t lut_n_right, lut_n_left;
};

struct K3Params {
    const uint64_t *cells;        // cell-major: cell j of read r at cells[j * stride + r]
    const uint32_t *lens;         // [n] stored length | N count << 16
    const uint16_t *samples;      // [n]
    scar_record *records;         // [n]
    const int32_t *sample_target; // [n_samples]
    const void *blob;             // static data staged in shared memory: tables | metas | phases | nmax
    uint32_t blob_bytes;          // multiple of 16
    uint32_t meta_off, phase_off, nmax_off;   // byte offsets into the blob (tables start at 0)
    int64_t n_reads;
    int64_t stride;               // reads per cell row
    int32_t W;                    // cells per read
    int32_t n_samples;
    int32_t do_phasing;
    int32_t min_length;
    uint64_t hr_code;             // donor, 2 bits/base
    int32_t hr_len;
    int32_t compact_tables;       // layout of the tables in the blob (ScarTableDesc)
    // ---- fused mode (scar_classify_raw_dev): raw read bytes in; packing (K1), classification (K3) and, when
    // `aggregate`, the per-sample accumulators and the scar table (K4) in ONE pass over a tile in shared memory
    int32_t platform;
    scar_ragged seq;              // raw reads (device pointers)
    uint32_t raw_cap;             // bytes of the tile's raw staging buffer (the packed cells reuse it)
    int32_t aggregate;            // bit 0: aggregate; bit 1 (test hook): key tags without the segment
    uint64_t first_index;         // global index of read 0 (first-seen order across batches and ranks)
    ScarTableSlot *slots; uint64_t slot_mask; uint32_t max_probe; int32_t hist_smem;
    uint32_t *claimed;            // claimed-slot list of the table
    unsigned long long *n_entries, *counters, *unassigned, *phase_hist;
    int32_t *status;              // bit 0: read longer than 16 W, bit 1: table full
    int32_t phase_bins, phase_joint;
};

struct K4Params {
    scar_ragged seq;             // raw bytes (device), for the ins / mh segment
    const scar_record *records;
    int64_t n_reads;
    uint64_t first_index;
    ScarTableSlot *slots;
    uint64_t slot_mask;          // capacity - 1
    uint32_t *claimed;           // claimed-slot list [capacity]
    unsigned long long *n_entries;
    unsigned long long *counters;   // [n_samples][SCAR_N_COUNTERS]
    unsigned long long *unassigned;
    unsigned long long *phase_hist; // [n_samples][2][SCAR_PHASE_BINS]
    int32_t *status;             // bit 1: table full
    int32_t n_samples;
    int32_t platform;
    int32_t use_smem;            // counters + phase histograms staged in shared memory
    uint32_t max_probe;
    int32_t weak_keys;           // test hook: key tags without the segment (see key_hashes)
    int32_t phase_joint;         // 1: joint [PB][PB] phase histogram in shared memory, 0: two [PB] marginals
    int32_t phase_bins;          // PB = max phases + 2 (none, phase 0 .. max-1, n/a)
};

struct K4VerifyParams {
    scar_ragged seq; const scar_record *records; int64_t n_reads; uint64_t first_index;
    const ScarTableSlot *slots; uint64_t slot_mask; unsigned long long *n_collisions; int32_t platform;
    uint32_t max_probe; int32_t weak_keys;
};

cudaError_t scar_launch_window(const K3Params &P, bool smem_tables, int sm_count, cudaStream_t stream, int *launches);
cudaError_t scar_window_plan(const K3Params &P, bool smem_tables, int *stage, int *per_sm, int *smem_bytes);
// fused K1+K3(+K4) from raw bytes; false when the problem is outside what the fused kernel takes (the caller then
// runs K1 -> K3 -> K4)
bool scar_fused_eligible(int W, int platform);
cudaError_t scar_launch_fused(K3Params P, bool smem_tables, int max_read_len, int sm_count, cudaStream_t stream, int *launches);
cudaError_t scar_fused_plan(K3Params P, bool smem_tables, int max_read_len, int *per_sm, int *smem_bytes);
cudaError_t scar_launch_pack(const scar_ragged &seq, int64_t n, int W, int platform, uint64_t *cells,
                             int64_t stride, uint32_t *lens, int32_t *status, int sm_count, cudaStream_t stream, int *launches);
cudaError_t scar_launch_demux(const K2Params &P, int sm_count, cudaStream_t stream, int *launches);
cudaError_t scar_launch_match_maker(const scar_ragged &q, const scar_ragged &u, int64_t n, int32_t *out,
                                    cudaStream_t stream);
cudaError_t scar_launch_aggregate(K4Params P, int sm_count, cudaStream_t stream, int *launches);
cudaError_t scar_launch_verify(const K4VerifyParams &P, int sm_count, cudaStream_t stream);
cudaError_t scar_launch_export(const ScarTableSlot *slots, const uint32_t *claimed, const unsigned long long *n_entries,
                               scar_entry *out, uint64_t out_capacity, int sm_count, cudaStream_t stream);
cudaError_t scar_launch_clear(ScarTableSlot *slots, const uint32_t *claimed, unsigned long long *n_entries, int sm_count,
                              cudaStream_t stream);
cudaError_t scar_launch_merge(const K4Params &P, const scar_entry *in, uint64_t n_in, int sm_count, cudaStream_t stream);
cudaError_t scar_launch_export_parts(const ScarTableSlot *slots, const uint32_t *claimed, const unsigned long long *n_entries,
                                     scar_entry *out, uint32_t n_parts, uint64_t region, int32_t *status, int sm_count,
                                     cudaStream_t stream);
cudaError_t scar_launch_homeology(const scar_ragged &cons, const scar_ragged &ins, const scar_ragged &targets,
                                  const scar_ragged &regions, const int32_t *clj, const int32_t *crj, const int32_t *tlj,
                                  const int32_t *trj, const int32_t *target_idx, int32_t n_targets, int64_t n, void *out,
                                  cudaStream_t stream);
cudaError_t scar_launch_alignment(const scar_ragged &genomic, const scar_ragged &consensus, const int32_t *cutposition,
                                  const int32_t *gap_con_len, const int32_t *gap_genome_len, int64_t n, int32_t max_spans,
                                  void *counts, int32_t *spans, int sm_count, cudaStream_t stream);
#define SCAR_MAX_READ_LEN 1024        // stored length of a read (16-bit fields of the records, nmax table)
#define SCAR_MAX_PEERS 32
cudaError_t scar_launch_export_p2p(const ScarTableSlot *slots, const uint32_t *claimed, const unsigned long long *n_entries,
                                   scar_entry *const *peer_base,
                                   uint32_t n_parts, uint32_t my_rank, uint64_t region, unsigned long long *cursors,
                                   int32_t *status, int sm_count, cudaStream_t stream);
cudaError_t scar_launch_merge_parts(const K4Params &P, const scar_entry *in, uint32_t n_parts, uint64_t region,
                                    int sm_count, cudaStream_t stream);

// GPU FASTQ indexer (scar_fastq_dev.cu); tiles are 4096 bytes
int64_t scar_fastq_offsets_elems(int64_t n_bytes);
cudaError_t scar_fastq_index_launch(const uint8_t *text, int64_t n_bytes, uint32_t *tile_counts, int64_t *tile_offsets,
                                    unsigned long long *total, cudaStream_t s);
cudaError_t scar_fastq_positions_launch(const uint8_t *text, int64_t n_bytes, const int64_t *tile_offsets, int64_t *nl_pos,
                                        int64_t nl_capacity, int64_t byte_base, cudaStream_t s);
cudaError_t scar_fastq_records_launch(const uint8_t *text, int64_t n_bytes, const int64_t *nl_pos, int64_t n_newlines,
                                      int64_t first_record, int64_t n_records, int want_fields, int64_t *seq_off, int32_t *seq_len,
                                      int64_t *f0_off, int32_t *f0_len, int64_t *f1_off, int32_t *f1_len, int32_t *status,
                                      int sm_count, cudaStream_t s, int32_t *max_len = nullptr);   // max_len: atomicMax of the sequence lengths
