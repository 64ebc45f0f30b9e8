/* kmx.h -- C ABI of libkmx_sm100.so, the B200 (sm_100a) engine for the kmtricks hot path
 * repart -> superk -> count -> merge (+ Bloom rows / bit transpose).
 *
 * The reference (tlemane/kmtricks v1.6.0) is single-process C++ with no FFI; its stages talk
 * through files in the run directory.  Each entry point below replaces the compute of one
 * reference task (include/kmtricks/task.hpp) -- the host keeps the CLI, run-dir layout, file
 * headers and IMergePlugin calls and hands the bytes to / takes the bytes from this library.
 * See INTEGRATION.md for the binding a kmtricks maintainer would add.
 *
 * Conventions: plain pointers and sizes, status-code returns (0 = KMX_OK), no exceptions or
 * STL across the boundary, one opaque context per device and host thread (one CUDA stream per
 * context), caller-owned host buffers, library-owned device buffers.  All multi-word k-mers
 * are little-endian 64-bit words, low word first (gatb LargeInt<2>, LargeInt2.pri:151-154).
 * There is NO CPU fallback: every call fails with KMX_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef KMX_H
#define KMX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct kmx_ctx kmx_ctx;

enum {
  KMX_OK = 0,
  KMX_ERR_ARG = 1,      /* bad argument / unsupported parameter */
  KMX_ERR_CUDA = 2,     /* CUDA runtime error or no device */
  KMX_ERR_FORMAT = 3,   /* input text is not strict 4-line FASTQ (use kmx_superk_push_reads) */
  KMX_ERR_NOMEM = 4,    /* device memory exhausted */
  KMX_ERR_STATE = 5     /* call order violated */
};

/* key kind -- reference "count_format" (include/kmtricks/cmd/cmd_common.hpp:67-100) */
enum { KMX_KEY_KMER = 0, KMX_KEY_HASH = 1 };

/* matrix row format -- reference MODE (cmd_common.hpp:102-118) */
enum {
  KMX_FMT_COUNT = 0,    /* [w x u64 key][N x u32]          merge.hpp:262-272 / :519-529 */
  KMX_FMT_PA = 1,       /* [w x u64 key][ceil(N/8) bytes]  merge.hpp:274-286 / :546-558 */
  KMX_FMT_BF = 2,       /* dense W x ceil(N/8) bytes        merge.hpp:575-600 (hash keys only) */
  KMX_FMT_BFT = 3       /* 8*ceil(N/8) rows x W/8 bytes     merge.hpp:631-644 (hash keys only) */
};

typedef struct {
  uint32_t kmer_size;        /* k, 8 <= k <= 63 here (w = ceil(k/32) words)                        */
  uint32_t minim_size;       /* m, 4 <= m <= 12, m <= k (table is uint16[4^m])                     */
  uint32_t nb_partitions;    /* P >= 1                                                             */
  uint32_t key_kind;         /* KMX_KEY_KMER | KMX_KEY_HASH                                        */
  uint64_t window_bits;      /* W = HashWindow::get_window_size_bits() (hash.hpp:31-38); hash only */
  const uint16_t* repart_table; /* [4^m] minimizer -> partition (Repartitor, PartiInfo.hpp:381;
                                   repartition.hpp:45-92).  Copied by kmx_create.                  */
  uint32_t nb_samples;       /* N: number of sample slots the context will hold                    */
  uint32_t reserved;
} kmx_params;

/* ---- lifetime ------------------------------------------------------------------------ */
int kmx_create(int device, const kmx_params* prm, kmx_ctx** out);
void kmx_destroy(kmx_ctx* ctx);
const char* kmx_last_error(const kmx_ctx* ctx);   /* thread-compatible; valid until next call */
/* number of kernels launched by this context since creation (bench.py "gpu_launches") */
uint64_t kmx_launch_count(const kmx_ctx* ctx);
int kmx_sync(kmx_ctx* ctx);
/* raw handle of the context's CUDA stream (cudaStream_t), so the host can time with events */
void* kmx_stream(kmx_ctx* ctx);

/* ---- stage 1: SuperKTask::exec (task.hpp:255-353) --------------------------------------
 * replaces KmFillPartitions<span>::operator() (gatb/fill_partitions.hpp:59-105) +
 * Sequence2SuperKmer (Sequence2SuperKmer.hpp:90-158) + Model::iterate (Model.hpp:725-765).
 * One sample at a time: begin, push any number of text blocks, end.                        */
int kmx_superk_begin(kmx_ctx* ctx);
/* `text`: a whole number of strict 4-line FASTQ records ('\n' or '\r\n' line ends).
 * on_device != 0: `text` is a device pointer (HBM-resident input).                          */
int kmx_superk_push_fastq(kmx_ctx* ctx, const char* text, size_t nbytes, int on_device);
/* host-parsed sequences (FASTA, multi-line, gz...): sequence i = seqs[off[i] .. off[i+1])     */
int kmx_superk_push_reads(kmx_ctx* ctx, const char* seqs, const uint64_t* off, size_t nseq);
/* finishes the sample; kmers_per_partition[P] = the numbers the reference writes to
 * partition_infos/<id>.pinfo (gatb_utils.hpp:46-51).  May be NULL.                          */
int kmx_superk_end(kmx_ctx* ctx, uint64_t* kmers_per_partition);

/* Repartition estimate on the device -- RepartTask (task.hpp:170-222) samples the banks on the CPU
 * (gatb RepartitionAlgorithm.cpp:395-492) to balance the partitions; here: while enabled, every stage-1 launch of this
 * context also adds each super-k-mer's k-mers to the load of its minimizer.  Run a sample of the input through
 * kmx_superk_* on a context created with any table, read the 4^m loads, assign minimizers to partitions (the host does
 * the longest-processing-time assignment, kmx_main.cpp: balanced_table) and create the real context with that table.
 * (A stage-1 launch that is redone after a bucket overflow counts twice: an estimate, not an exact census.)             */
int kmx_minimizer_load_enable(kmx_ctx* ctx, int on);
int kmx_minimizer_load_get(kmx_ctx* ctx, uint64_t* load /* [4^m] */);

/* ---- stage 2: CountTask / HashCountTask::exec (task.hpp:367-495) ------------------------
 * replaces KmerPartCounter / HashPartCounter::execute (gatb/sorting_count.hpp:637-650,934-943)
 * + Kmer/HashCountProcessor (gatb/count_processor.hpp:61-70,135-146): for every partition of
 * the sample just finished by kmx_superk_end, the ascending list of (key, count >= hard_min),
 * counts saturating at 2^32-1.  The lists stay in HBM under slot `sample`.                   */
int kmx_count_sample(kmx_ctx* ctx, uint32_t sample, uint32_t hard_min);
/* Stage 1 + 2 for many samples in one call (what TaskScheduler::exec_superk_count,
 * task_scheduler.hpp:251-348, does with its thread pool): sample i is the strict 4-line FASTQ
 * block texts[i] (nbytes[i] bytes, host or device per on_device), counted with hard_min[i]
 * into slot sample_ids[i] (NULL = i).  `nlanes` (1..8) samples are in flight at once, each on
 * its own CUDA stream, so host->device copies and small read-backs overlap other samples'
 * kernels.  kmers_per_partition (n*P, may be NULL) receives every sample's .pinfo vector.     */
int kmx_run_samples(kmx_ctx* ctx, uint32_t n, const char* const* texts, const size_t* nbytes, int on_device,
                    const uint32_t* sample_ids, const uint32_t* hard_min, uint32_t nlanes,
                    uint64_t* kmers_per_partition);
/* Lane-addressed variants for a host that STREAMS its inputs (what TaskScheduler does with one task per pool thread,
 * task_scheduler.hpp:251-348): a lane is an independent in-flight sample (own CUDA stream, staging and buckets); one host
 * thread drives one lane, several lanes run concurrently, so reading / inflating the next block of one sample overlaps the
 * device work of the others and no input ever has to be resident as a whole.  kmx_lanes makes lanes 0..n-1 exist (n <= 8).
 * A text block must be a whole number of FASTQ records and smaller than 4 GiB: push a large file in several blocks.     */
int kmx_lanes(kmx_ctx* ctx, uint32_t n);
int kmx_lane_superk_begin(kmx_ctx* ctx, uint32_t lane);
int kmx_lane_superk_push_fastq(kmx_ctx* ctx, uint32_t lane, const char* text, size_t nbytes, int on_device);
int kmx_lane_superk_push_reads(kmx_ctx* ctx, uint32_t lane, const char* seqs, const uint64_t* off, size_t nseq);
int kmx_lane_superk_end(kmx_ctx* ctx, uint32_t lane, uint64_t* kmers_per_partition);
int kmx_lane_count_sample(kmx_ctx* ctx, uint32_t lane, uint32_t sample, uint32_t hard_min);
/* kmx_lane_count_sample + the sample's abundance histogram (--hist: KHist, histogram.hpp:34-207; every distinct key of the
 * sample is binned BEFORE the hard-min test, count_processor.hpp:61-70,135-146).
 * out = [uniq, total, oob_lower_unique, oob_lower_total, oob_upper_unique, oob_upper_total,
 *        hist_unique[upper-lower+1], hist_total[upper-lower+1]]  (the reference uses lower 1, upper 255).               */
int kmx_lane_count_sample_hist(kmx_ctx* ctx, uint32_t lane, uint32_t sample, uint32_t hard_min, uint32_t lower, uint32_t upper,
                               uint64_t* out);
/* size of / copy out one list == body of counts/partition_P/<id>.kmer|.hash
 * keys: n*w u64 (w = 1 for hash keys), counts: n u32.                                        */
int kmx_counts_size(kmx_ctx* ctx, uint32_t sample, uint32_t partition, uint64_t* n);
int kmx_counts_get(kmx_ctx* ctx, uint32_t sample, uint32_t partition, uint64_t* keys, uint32_t* counts);
/* inject a list read from a counts/ file (interop with a reference-produced run-dir)        */
int kmx_counts_put(kmx_ctx* ctx, uint32_t sample, uint32_t partition, const uint64_t* keys,
                   const uint32_t* counts, uint64_t n);
/* per-sample Bloom window of one partition: HashVecProcessor (count_processor.hpp:84-120),
 * W/8 bytes, bit (h - W*p) LSB-first.                                                       */
int kmx_counts_vector(kmx_ctx* ctx, uint32_t sample, uint32_t partition, uint8_t* bits);

/* ---- stage 3/4: KmerMergeTask / HashMergeTask::exec (task.hpp:690-870) ------------------
 * replaces KmerMerger / HashMerger::next + write_as_{bin,pa,bf,bft} (merge.hpp:183-286,
 * 441-644) and MergeStatistics (merge.hpp:49-100).                                          */
typedef struct {
  const uint32_t* soft_min;   /* [N] per-sample abundance min (--soft-min / fof)              */
  uint32_t recurrence_min;    /* --recurrence-min                                            */
  uint32_t share_min;         /* --share-min ("save_if"), 0 = no rescue                      */
  uint32_t format;            /* KMX_FMT_*                                                   */
  uint32_t emit_all;          /* !=0: emit every merged row (plugin loaded);
                                 COUNT format only; row_keep tells the default decision      */
} kmx_merge_params;

typedef struct {
  uint64_t n_rows;            /* rows emitted (COUNT/PA) ; W for BF ; 8*ceil(N/8) for BFT     */
  uint64_t row_bytes;         /* bytes per emitted row                                       */
  uint64_t n_union;           /* distinct keys seen before the keep filter                   */
} kmx_merge_result;

/* merges partition `partition` over all N sample slots; the body bytes stay on the device */
int kmx_merge_partition(kmx_ctx* ctx, uint32_t partition, const kmx_merge_params* mp, kmx_merge_result* res);
/* copy out: body = n_rows*row_bytes bytes (exactly the bytes after the file header of
 * matrices/matrix_P.<ext>), stats = 6*N u64 in merge_info order (merge.hpp:72-83),
 * row_keep = n_rows bytes (emit_all only).  Any pointer may be NULL.                         */
int kmx_merge_get(kmx_ctx* ctx, void* body, uint64_t* stats, uint8_t* row_keep);
/* device pointer to the last merge body (valid until the next merge on this context)        */
const void* kmx_merge_body_device(kmx_ctx* ctx);

/* ---- stage 4 stand-alone: BitMatrix::transpose (bitmatrix.hpp:209-214,238-289) ----------
 * in: nrows x ncols bits, rows of ncols/8 bytes, LSB-first; out: ncols x nrows bits.
 * nrows and ncols must be multiples of 8.  Host buffers.                                    */
int kmx_transpose_bits(kmx_ctx* ctx, const uint8_t* in, uint64_t nrows, uint64_t ncols, uint8_t* out);

/* ---- multi-GPU (one process per GPU) ----------------------------------------
 * Stage 1 shards over samples, stages 2-4 over partitions: rank g owns the contiguous block
 * [g*P/world, (g+1)*P/world).  kmx_dist_run_samples = kmx_run_samples + ONE exchange per
 * sample: the bucket regions of every partition travel to the partition's owner (grouped
 * ncclSend/ncclRecv over NVLink), who counts them.  Rank r's local sample i lands in slot
 * r*n_local + i on every rank (for the partitions that rank owns), so nb_samples must be
 * >= world*n_local.  Afterwards each rank merges its own partitions with kmx_merge_partition.
 * Rendezvous: rank 0 draws `nlanes` ids with kmx_dist_unique_id, the host broadcasts them
 * (torch.distributed / MPI / files), every rank calls kmx_dist_init with all of them.        */
int kmx_dist_unique_id(uint8_t* out128);
int kmx_dist_init(kmx_ctx* ctx, int rank, int world, uint32_t nlanes, const uint8_t* ids /* nlanes*128 */);
int kmx_dist_owner(const kmx_ctx* ctx, uint32_t partition, int world);
int kmx_dist_run_samples(kmx_ctx* ctx, uint32_t n_local, const char* const* texts, const size_t* nbytes, int on_device,
                         const uint32_t* hard_min, uint64_t* kmers_per_partition);
/* The same for one BATCH of a longer run (inputs streamed batch by batch, or more samples than fit HBM as text): rank r's
 * sample i of this batch lands in slot r*n_local_total + slot_base + i.  Every rank calls it with the same n, slot_base
 * and n_local_total; an empty text (nbytes 0) is a valid sample (pads the last batch).                                   */
int kmx_dist_run_batch(kmx_ctx* ctx, uint32_t n, const char* const* texts, const size_t* nbytes, int on_device,
                       const uint32_t* hard_min, uint32_t slot_base, uint32_t n_local_total, uint64_t* kmers_per_partition);
/* how many of the communicators' lanes kmx_dist_run_samples uses (1..nlanes of kmx_dist_init; every rank the same
 * value).  1 = samples strictly one after the other on one stream: the per-kernel profile spans are then exclusive.     */
int kmx_dist_set_lanes(kmx_ctx* ctx, uint32_t nlanes);

/* ---- utilities (benchmark / tests) ------------------------------------------------------ */
/* device twin of kmtricks_b200/synth.py: writes R records of 2L+15 bytes to dev_out        */
int kmx_synth_fastq(kmx_ctx* ctx, uint64_t seed, uint32_t sample, uint64_t first_read, uint64_t R,
                    uint32_t L, uint64_t G, double d, double e, int revcomp, char* dev_out);
int kmx_dev_alloc(kmx_ctx* ctx, size_t nbytes, void** dev_ptr);
int kmx_dev_free(kmx_ctx* ctx, void* dev_ptr);
int kmx_memcpy_d2h(kmx_ctx* ctx, void* host, const void* dev, size_t nbytes);
int kmx_memcpy_h2d(kmx_ctx* ctx, void* dev, const void* host, size_t nbytes);
/* pinned host memory for the end-to-end path */
int kmx_host_alloc(size_t nbytes, void** host_ptr);
int kmx_host_free(void* host_ptr);
/* the next KMX_FMT_BF / KMX_FMT_BFT merges write their body to this device buffer (>= body
 * bytes + 8) instead of the context's own; NULL restores the default                         */
int kmx_set_merge_output(kmx_ctx* ctx, void* dev_ptr, size_t cap_bytes);
/* per-kernel-class device timing with CUDA events on the context's stream (bench roofline) */
enum { KMX_PROF_INDEX = 0, KMX_PROF_S1 = 1, KMX_PROF_HASH_HIST = 2, KMX_PROF_HASH_EMIT = 3, KMX_PROF_EXPAND = 4,
       KMX_PROF_SORT = 5, KMX_PROF_RLE = 6, KMX_PROF_MERGE = 7, KMX_PROF_TRANSPOSE = 8, KMX_PROF_FILL = 9,
       KMX_PROF_EXCHANGE = 10, KMX_PROF_KINDS = 11 };
int kmx_profile_enable(kmx_ctx* ctx, int on);
int kmx_profile_reset(kmx_ctx* ctx);
int kmx_profile_get(kmx_ctx* ctx, int kind, double* total_ms, uint64_t* count);
/* drops all per-sample lists and buckets (keeps parameters) */
int kmx_reset(kmx_ctx* ctx);
/* bytes of device memory currently held by the context */
uint64_t kmx_device_bytes(const kmx_ctx* ctx);
/* path counters (tests, bench): FASTQ blocks whose stage 1 found its reads in the newline masks itself
 * (no line-index pass) / blocks that went through the line index / hash-count passes on the binned path */
enum { KMX_STAT_S1_SELF_INDEXED = 0, KMX_STAT_S1_INDEXED = 1,
       KMX_STAT_HASH_BINNED = 2,   /* samples counted by the binned shared-memory path (hash keys, k <= 32) */
       KMX_STAT_S1_RETRY = 3,      /* stage-1 launches redone because a bucket region was too small */
       KMX_STAT_EXCH_BYTES = 4,    /* bytes this rank sent through the bucket exchange (kmx_dist_run_samples) */
       KMX_STAT_KINDS = 5 };
uint64_t kmx_stat(const kmx_ctx* ctx, int which);

#ifdef __cplusplus
}
#endif
#endif /* KMX_H */
