// kmx_internal.h -- declarations shared by the .cu files of libkmx_sm100 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace kmx {

typedef unsigned long long u64;
typedef unsigned int u32;


bool kmx_env_flag(const char* name);       // A/B switches from the environment (s2_hash.cu)

// ---- stage 1 (s1_superk.cu) -------------------------------------------------------------
struct S1Args {
  const uint8_t* text;          // text base
  u64 text_bytes;
  const u32* seg_start;         // per segment: offset in text
  const u32* seg_len;           // per segment: number of bases
  u64 nseg;
  int k, m, wlen;               // wlen = k - m + 1
  int max_nk;                   // max k-mers per record
  u32 P;
  const uint16_t* repart;       // [4^m]
  void* records;                // bucket slab (uint4 units)
  const u64* boff;              // [P] record offset of each partition's region
  const u32* bcap;              // [P] capacity (records)
  u32* cursor;                  // [P] records appended (may exceed bcap -> overflow)
  u64* kcnt;                    // [P] k-mers appended
  u32* overflow;                // set when a record did not fit
  u32 pack_words;               // 32-bit words of packed bases kept per segment = ceil(max segment length / 16)
  u32 stage_cap;                // staging capacity (cut events)
  u32 flush_thr;                // flush when staged > flush_thr
  u64* mload;                   // NULL, or [4^m]: every record adds its k-mers to the load of its minimizer (repartition estimate)
};
size_t s1_smem_bytes(u32 pack_words, u32 stage_cap, int wlen, u32 P);
u64 fq_num_tiles(const uint8_t* text, u64 nbytes);
cudaError_t launch_fq_index(const uint8_t* text, u64 nbytes, u32* tile_counts, u64* tile_prefix, u64* nlmask /* [tiles * 256] */,
                            u64* d_total, u32* seq_start, u32* seq_len, u64 nrec_cap, u32* flags,
                            int phase, cudaStream_t st, u64* launches);
cudaError_t launch_s1(int W, const S1Args& a, cudaStream_t st, u64* launches);
// position-parallel variant for short reads (s1_v5.cu)
namespace s1v5 { struct Geo; }
bool s1_v5_usable(u32 max_len, int k, int m, u32 P, s1v5::Geo* geo, size_t* smem);
// idx == NULL: reads given by a.seg_start / a.seg_len.  Otherwise the kernel finds its own reads in the
// newline masks of the FASTQ count pass (no line-index pass, no seq_start / seq_len arrays): a.nseg = records
static const u32 FQ_TILE_BYTES = 16384;        // text bytes per tile of the count pass (256 threads x 64 bytes)
static const u32 S1_FUSED_R = 32;              // reads per CTA of the self-indexing launch
struct S1Idx {
  const u64* nlmask64;        // one bit per text byte, whole tiles, in the coordinates of the 16-byte-aligned base below the text
  const u32* cta_pos;         // [ceil(nrec / S1_FUSED_R)] position (mask coordinates) of newline number 4 * S1_FUSED_R * cta
  u64 ntiles, lead, tot;      // tot = lead + text bytes
  u32* flags;                 // [0] not strict 4-line FASTQ, [3] a read is longer than the launch geometry
  u32 geo_maxlen;
};
cudaError_t launch_s1_v5(int W, const S1Args& a, const s1v5::Geo& geo, size_t smem, const S1Idx* idx, cudaStream_t st, u64* launches);
cudaError_t launch_fq_cta_pos(const u64* nlmask64, const u64* tile_prefix, u64 ntiles, u32* cta_pos, u64 ncta, cudaStream_t st, u64* launches);

// ---- stage 2 (s2_count.cu) --------------------------------------------------------------
struct S2Common {
  int W;                  // words per k-mer (1 or 2)
  int k;
  u32 P;
  const void* records;    // bucket slab
  const u64* boff;        // [P] device
  const u32* bcnt;        // [P] device: records per partition
  const u64* kcnt;        // [P] device: k-mers per partition (hash path, 16-bit counters: overflow check)
  u32 max_bcnt;           // host copy of max(bcnt)
};

// hash keys, histogram path: partitions are processed in groups of gp windows
// (phase 0: RED fill; phase 1: single-pass ordered sweep = chained scan + emit + re-zero)
static const u32 HIST_SUB = 8192;    // slots per chunk of the sweep (one CTA, 8 warps x 1024 slots)
u32 hash_sweep_chunks_per_window(u64 Wbits);
// staging of the sweep: per 1024-slot slice a run of (slot offset, count) in slot order
struct SweepStage { uint16_t* idx; u32* cnt; u32* slice_counts; u32* done /* [gp] tiles done per window | ticket */; u64* win_sum /* [gp] */; };
cudaError_t launch_hash_group(const S2Common& c, u64 Wbits, u64 mod_d, u64 mod_mlo, u64 mod_mhi, u32* hist, u32 hard_min,
                              u32 p0, u32 gp, u32 group_idx, u32* chunk_counts, u64* chunk_off, SweepStage stage, u64* list_off, u64* meta, u32* flags,
                              u64* out_keys, u32* out_counts, const u32* win_part, cudaStream_t st, u64* launches, int phase, bool h16);
cudaError_t launch_scan_u32(const u32* in, u64* out, u64 n, u64* total, cudaStream_t st, u64* launches);

// hash keys, k <= 32: binned shared-memory counting (s2_bin.cu).  "Window" v = one (sample, partition) bucket region.
struct HashBinArgs {
  const void* records; const u64* boff; const u32* bcnt;   // [nwin] device: bucket region of every window
  int k; u32 nwin;
  u64 Wbits; u32 NB, bs_log;      // slots per window, bins per window, log2(slots per bin)
  const u32* tile_pref;           // [nwin+1] device: tiles of hash_bin_tile_records() records before window v
  const u64* win_base;            // [nwin] device: first bin-buffer entry of window v (multiple of 8)
  const u32* win_cap;             // [nwin] device: entries per bin of window v (multiple of 8)
  u32* bin_cursor;                // [nwin*NB] zeroed: entries appended per bin (keeps counting past win_cap)
  uint16_t* binbuf;               // 16-bit slot offsets, grouped by (window, bin)
  u64* status;                    // [nwin*NB] zeroed: look-back words of pass B
  u32* tickets;                   // [2] zeroed
  u32* flags;                     // [0] output space ran out [1] a 16-bit counter wrapped [2] a bin region overflowed
  u64* list_off;                  // [nwin] out: first output entry of window v
  u64* meta;                      // [0] out: total entries, [2] in: output capacity
  u64* out_keys; u32* out_counts;
  const u32* win_part;            // NULL: window v holds partition v
  u32 hard_min;
};
u32 hash_bin_max_bins();
u32 hash_bin_tile_records();
cudaError_t launch_hash_binned(const HashBinArgs& a, u32 total_tiles, int phase, bool h16, cudaStream_t st, u64* launches);

// generic path: expand -> keys, segmented radix sort, run-length
cudaError_t launch_expand_keys(const S2Common& c, int key_kind, u64 Wbits, u64 mod_d, u64 mod_mlo, u64 mod_mhi,
                               const u64* koff /* [P] device: key offset of partition */,
                               u32* kcursor /* [P] device, zeroed */, u64* keys_lo, u64* keys_hi,
                               cudaStream_t st, u64* launches);

// ---- stage 3/4 (s3_merge.cu) ------------------------------------------------------------
struct MergeList { const u64* lo; const u64* hi; const u32* cnt; u64 n; };
// dense (hash bf/bft) path: rows addressed by key - lower
cudaError_t launch_dense_solid(const MergeList* d_lists, u32 N, const u32* d_soft, u64 lower, u32* solid_in,
                               u64 max_n, cudaStream_t st, u64* launches);
cudaError_t launch_dense_emit(const MergeList* d_lists, u32 N, const u32* d_soft, u32 rmin, u32 share,
                              u64 lower, const u32* solid_in /* may be NULL when not needed */,
                              uint8_t* slab, u32 row_bytes, u64* stats /* 6N */, u64 max_n,
                              cudaStream_t st, u64* launches);
cudaError_t launch_row_keep(const u32* solid_in, u64 nrows, u32 rmin, u32 emit_all, u32* keep_flag, cudaStream_t st, u64* launches);

// ---- stage 4 (s4_bits.cu) ---------------------------------------------------------------
cudaError_t launch_transpose_bits(const uint8_t* in, u64 nrows, u64 ncols, uint8_t* out, cudaStream_t st, u64* launches);
cudaError_t launch_hash_vector(const u64* keys, u64 n, u64 lower, uint8_t* bits, cudaStream_t st, u64* launches);

// ---- synth (synth.cu) -------------------------------------------------------------------
cudaError_t launch_synth_fastq(u64 seed, u32 sample, u64 first_read, u64 R, u32 L, u64 G, u32 thr_d, u32 thr_e,
                               int revcomp, char* out, cudaStream_t st, u64* launches);

}  // namespace kmx
