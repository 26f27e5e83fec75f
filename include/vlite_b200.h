/* vlite_b200 -- C ABI of the B200-native vlite retrieval hot path.
 *
 * The reference (sdan/vlite) has no FFI: its seam is three Python call sites
 * (SURVEY.md section 8b).  Each entry point below names the reference code it
 * replaces (paths relative to the reference checkout).  Plain pointers and
 * sizes only; no torch / C++ types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success and a negative VL_E* code on
 *     failure; vl_last_error() then returns a thread-local message.
 *   - corpus rows and queries are RAW UBINARY bytes (np.packbits output,
 *     MSB-first within a byte), width W in {32, 64, 128} bytes per row (256 / 512 / 1024 bits).  The
 *     reference's stored encoding int8(u8)-128 (vlite/model.py:44) gives
 *     identical XOR scores; vl_corpus_append_f32 converts it on the device.
 *   - pointers named *_any may be host or device memory (the library asks
 *     the CUDA runtime); *_dev must be device memory on the handle's device.
 *   - one CUDA stream per handle (vl_corpus_set_stream); calls on one handle
 *     are serialised by the caller, different handles are independent --
 *     the same contract the reference's single mutable VLite object has
 *     (vlite/server.py:14).  When every pointer passed to a call is device
 *     memory the call only ENQUEUES work on the handle's stream and returns
 *     without synchronising -- for every k: the exact fallback pass of k > 32
 *     is enqueued behind a device-side gate instead of being decided by a host
 *     read-back; with any host pointer it returns after the data has landed.
 *   - results are ordered ascending by (score, global row); slots past the
 *     number of eligible rows hold VL_SENTINEL_SCORE / VL_SENTINEL_ROW.
 *   - there is NO CPU fallback: without a usable sm_100 device every compute
 *     entry point fails with VL_ENODEV.
 */
#ifndef VLITE_B200_H
#define VLITE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VL_OK        0
#define VL_EINVAL   -1   /* bad argument */
#define VL_ECUDA    -2   /* CUDA runtime error (message in vl_last_error) */
#define VL_ENOMEM   -3   /* device or host allocation failed */
#define VL_ENODEV   -4   /* no usable sm_100 device */
#define VL_EIO      -5   /* file error in vl_corpus_load_file */
#define VL_ERANGE   -6   /* stored value outside the encodable range */

#define VL_SENTINEL_SCORE 0xFFFFFFFFu
#define VL_SENTINEL_ROW   0xFFFFFFFFFFFFFFFFull

/* score function */
#define VL_METRIC_HAMMING 0  /* popcount(q ^ e): vlite/model.py:71-74 hamming_distance,
                                vlite/index.py:30 (on unpacked bits) */
#define VL_METRIC_XORSUM  1  /* sum_j (q_j ^ e_j) on byte values: the live
                                torch.sum(q.int() ^ E.int(), dim=1), vlite/model.py:82 */

/* source encodings for vl_corpus_append_f32 / vl_corpus_load_file */
#define VL_SRC_U8          0 /* raw ubinary bytes, W per row */
#define VL_SRC_F32_OFFSET  1 /* W float32 per row holding int8(u8)-128 in [-256,-1]
                                (vlite/model.py:44 as stored by vlite/ctx.py:70), or raw [0,255] */
#define VL_SRC_F32_BITS    2 /* 8*W float32 per row holding 0.0/1.0, one per bit (the legacy
                                layout of tests/contexts/vlite-bench.ctx) */

typedef struct vl_corpus vl_corpus;

/* ---- library ---------------------------------------------------------- */
const char* vl_version(void);
const char* vl_last_error(void);
/* number of CUDA devices of compute capability 10.x visible to the process */
int vl_device_count(void);
/* how many of this library's kernels the calling process has launched so far
 * (bench.py's "gpu_launches") */
uint64_t vl_launch_count(void);

/* ---- corpus residency: replaces the per-query rebuild of the (N, W) array
 * from the Python dict (vlite/main.py:143-146) and the dict appends of
 * VLite.add / set_batch (vlite/main.py:89-95, :264-268) ------------------- */
/* capacity_rows: rows to back with memory up front (0 = grow on demand); a shard holds at most
 * 2^31 - 1024 rows (local rows are 32-bit inside the kernels; global rows are 64-bit) */
int vl_corpus_create(int device, int width_bytes, uint64_t capacity_rows, vl_corpus** out);
int vl_corpus_destroy(vl_corpus* c);
/* stream: a cudaStream_t (as void*); NULL = the legacy default stream */
int vl_corpus_set_stream(vl_corpus* c, void* cuda_stream);
/* back at least capacity_rows rows with physical memory now (exact, e.g. a 16 GB shard) */
int vl_corpus_reserve(vl_corpus* c, uint64_t capacity_rows);
/* append n rows of raw ubinary bytes IN PLACE: the shard lives in a virtual address range
 * reserved once; when it fills, more physical HBM is mapped behind it (no copy, the base pointer
 * never moves).  *first_row_out = local index of the first appended row */
int vl_corpus_append(vl_corpus* c, const uint8_t* rows_any, uint64_t n, uint64_t* first_row_out);
/* append n rows given as float32 in one of the VL_SRC_F32_* encodings; the
 * f32 -> ubinary repack runs on the device.  replaces the row-by-row
 * struct.unpack of vlite/ctx.py:116-119 + np.array() of vlite/main.py:50 */
int vl_corpus_append_f32(vl_corpus* c, const float* vals_any, uint64_t n, int src_kind,
                         uint64_t* first_row_out);
/* stream rows straight from a byte range of a file (a CTX EMBEDDINGS payload,
 * vlite/ctx.py:112-119) through pinned staging into HBM; src_row_bytes is the
 * on-disk size of one row in that encoding (e.g. 256 for 64 float32).
 * max_rows caps how many rows are taken (0 = all that fit in byte_len) */
int vl_corpus_load_file(vl_corpus* c, const char* path, uint64_t byte_off, uint64_t byte_len,
                        int src_kind, uint64_t max_rows, uint64_t* first_row_out,
                        uint64_t* n_rows_out);
/* append n rows whose bits are the signs (x > 0, MSB-first like np.packbits) of an int8 document
 * store already on the device: docs_i8_dev is n x (8*W) int8, 16 B aligned.  Keeps the binary
 * corpus and the rescore documents consistent without a host round trip (config 3). */
int vl_corpus_append_sign_i8(vl_corpus* c, const int8_t* docs_i8_dev, uint64_t n, uint64_t* first_row_out);
/* overwrite one existing row (VLite.update with a new vector) */
int vl_corpus_set_row(vl_corpus* c, uint64_t row, const uint8_t* row_any);
/* delete = clear the row's bit in the live mask (replaces `del self.index[id]`,
 * vlite/main.py:198); rows keep their indices */
int vl_corpus_tombstone(vl_corpus* c, const uint64_t* rows_host, uint64_t n);
int vl_corpus_clear(vl_corpus* c);                       /* VLite.clear, vlite/main.py:296 */
int vl_corpus_rows(const vl_corpus* c, uint64_t* rows, uint64_t* live_rows);
int vl_corpus_width(const vl_corpus* c);
/* global index of local row 0 (row-sharding across GPUs) */
int vl_corpus_set_base_row(vl_corpus* c, uint64_t base_row);
/* copy rows [first, first+n) back to the host (CTX save, tests) */
int vl_corpus_read(vl_corpus* c, uint64_t first, uint64_t n, uint8_t* out_host);
/* device address of row 0 and of the live mask (borrowed; stable for the life of the handle) */
int vl_corpus_device_ptrs(vl_corpus* c, void** rows_dev, void** live_mask_dev);
/* fill rows [first, first+n) on the device with the counter-based synthetic
 * generator of oracle/synth.py (period = 0: all rows distinct); rows past the
 * current count are appended.  Bench / test data only. */
int vl_corpus_fill_synthetic(vl_corpus* c, uint64_t seed, uint64_t first, uint64_t n,
                             uint64_t period);
/* build a filter mask on the device from the synthetic per-row tags:
 * bit(row) = tag(seed, base_row + row) < threshold.  mask_dev holds
 * ceil(n/32) uint32 words.  Bench / test data only. */
int vl_synth_filter_mask(vl_corpus* c, uint64_t seed, uint64_t threshold, uint32_t* mask_dev);

/* ---- search: replaces EmbeddingModel.search (vlite/model.py:76-90) and the
 * row -> id step of VLite.rank_and_filter (vlite/main.py:150-156) ---------- */
/* queries_any: Q x W raw ubinary.  filter_mask_any: optional bit per local row
 * (bit row&31 of word row>>5; 1 = eligible), ANDed with the live mask INSIDE
 * the scan -- a true pre-filter, unlike the reference's post-filter over
 * top_k*4 candidates (vlite/main.py:159-165).  out_score_any: Q x k uint32,
 * out_row_any: Q x k uint64 global rows (base_row + local).  1 <= k <= VL_MAX_K */
#define VL_MAX_K 2048
int vl_search(vl_corpus* c, const uint8_t* queries_any, int n_queries, int k, int metric,
              const uint32_t* filter_mask_any, uint32_t* out_score_any, uint64_t* out_row_any);
/* the same search for queries given as FLOAT embeddings (n_queries x dims float32, host or device):
 * the quantisation tail of EmbeddingModel.embed (vlite/model.py:39-42: first 8*W features, x > 0,
 * np.packbits MSB-first) runs on the device in front of the scan, so VLite.retrieve
 * (vlite/main.py:113-117) is ONE call with one copy in and one copy out.  dims >= 8 * W. */
int vl_search_embeddings(vl_corpus* c, const float* x_any, int n_queries, int dims, int k, int metric,
                         const uint32_t* filter_mask_any, uint32_t* out_score_any,
                         uint64_t* out_row_any);
/* full score vector of ONE query over rows [first, first+n) (parity tests) */
int vl_scores(vl_corpus* c, const uint8_t* query_any, int metric, uint64_t first, uint64_t n,
              uint32_t* out_any);
/* merge n_lists candidate lists (each Q x k_in, ascending (score,row), sentinel
 * padded) into Q x k_out: the step after the all-gather of per-shard top-k
 * (SURVEY.md section 8e).  All pointers device memory.  List l starts at
 * scores_dev + l*score_list_stride and rows_dev + l*row_list_stride (strides in
 * ELEMENTS; 0 = dense, i.e. Q*k_in), so one all-gathered byte buffer holding
 * every rank's [scores | rows] block can be merged in place. */
int vl_merge_topk(int device, void* cuda_stream, const uint32_t* scores_dev, const uint64_t* rows_dev,
                  int n_lists, int n_queries, int k_in, int k_out,
                  uint64_t score_list_stride, uint64_t row_list_stride,
                  uint32_t* out_score_dev, uint64_t* out_row_dev);

/* rewrite LOCAL rows (as a shard with base_row 0 returns them) into GLOBAL rows for a shard made of
 * several appended segments: segment i holds local rows [seg_local[i], seg_local[i+1]) and global
 * rows seg_global[i] + (local - seg_local[i]).  Tables ascend; all pointers device memory.
 * (Row-sharded collections: appends go to the least-full shard with the global base fixed at append
 * time; the reference has no counterpart.) */
int vl_map_rows(int device, void* cuda_stream, uint64_t* rows_dev, uint64_t n,
                const uint64_t* seg_local_dev, const uint64_t* seg_global_dev, int n_segments);

/* ---- rescore (north-star piece 3; no reference implementation -- nearest
 * code is the unused cos_sim, vlite/utils.py:205-208) ----------------------- */
#define VL_QTYPE_F32 0
#define VL_QTYPE_I8  1
#define VL_QTYPE_GATHER_PROBE 2  /* measurement only: int8-shaped queries, same gathers / sort / stores, no dot
                                    products -- the HBM gather floor of a rescore (scores are meaningless) */
/* docs_i8_dev: n_docs x dims int8, row-major, indexed by LOCAL row.
 * queries_any: Q x dims float32 or int8.  cand_rows_any: Q x C uint64 GLOBAL
 * rows as vl_search wrote them (sentinels skipped).  Output: top k by
 * DESCENDING dot product, ties by ascending row; padded with -inf / sentinel. */
int vl_rescore(vl_corpus* c, const int8_t* docs_i8_dev, uint64_t n_docs, int dims,
               const void* queries_any, int qtype, const uint64_t* cand_rows_any,
               int n_queries, int n_cand, int k, float* out_score_any, uint64_t* out_row_any);
/* sign-quantise float32 embeddings to raw ubinary on the device:
 * (x > 0) -> packbits MSB-first (vlite/model.py:41-44 without the -128).
 * x_any: n x dims float32, out_any: n x dims/8 bytes. */
int vl_quantize_binary(int device, void* cuda_stream, const float* x_any, uint64_t n, int dims,
                       uint8_t* out_any);
/* The whole quantisation tail of EmbeddingModel.embed (vlite/model.py:34-44) in one launch, on
 * device memory only, so an encoder's output never round-trips through the host:
 * x_dev: n x dims float32 (the CLS embedding, normalised or not).  Each row is L2-normalised over
 * all `dims` features (x / max(||x||, 1e-12), model.py:35); the first `binary_dims` features are
 * sign-packed MSB-first into out_u8_dev (n x binary_dims/8 -- bit-identical to
 * vl_quantize_binary on the same columns); out_f32_dev (optional, n x dims) receives the
 * normalised row and out_i8_dev (optional, n x dims) rint(clamp(x_norm * i8_scale, -127, 127)):
 * the query / document forms vl_rescore takes.  Asynchronous on cuda_stream. */
int vl_quantize_embeddings(int device, void* cuda_stream, const float* x_dev, uint64_t n, int dims,
                           int binary_dims, float i8_scale, uint8_t* out_u8_dev, int8_t* out_i8_dev,
                           float* out_f32_dev);
/* bench data: fill n x dims int8 documents with the synthetic generator */
int vl_synth_docs_i8(int device, void* cuda_stream, uint64_t seed, uint64_t first_row, uint64_t n,
                     int dims, int8_t* out_dev);

/* ---- measurement helpers ---------------------------------------------- */
/* integer-pipe micro-benchmarks: kind 0 = POPC, 1 = LOP3, 2 = IADD3, 3 = IDP4A,
 * 4 = POPC+LOP3+IADD mix.  Runs `iters` dependent-chain steps per thread on a
 * full grid; returns lane-operations per second. */
int vl_microbench(int device, int kind, int iters, double* ops_per_sec_out, double* ms_out);
/* wall time of the most recent vl_search / vl_rescore on this handle, measured
 * with CUDA events on the handle's stream around the kernels only (no copies);
 * scan_ms = the scan kernel alone */
int vl_last_timing(vl_corpus* c, float* total_ms, float* scan_ms);
int vl_set_timing(vl_corpus* c, int enabled);
/* tuning knobs for experiments and tests (0 = automatic): row_splits = CTAs per query tile;
 * stages = depth of the TMA ring, or -1 to force the exact multi-pass path for k > 32 */
int vl_set_tuning(vl_corpus* c, int row_splits, int stages);
/* debug build only (-DVL_DEBUG_BOUNDS, `python -m vlite_b200.build --debug`): how many device-side
 * bounds / protocol assertions failed on `device` so far and the id of the first (0 = none).  The
 * shipped build returns VL_EINVAL.  Stands in for compute-sanitizer, which the GPU pool keeps closed. */
int vl_debug_checks(int device, uint32_t* failures_out, uint32_t* first_id_out);
/* which scan kernel answers vl_search on this handle.  0 = automatic (default): batches of at least 16
 * queries (XORSUM on shards over ~1M rows: 44, 40 at W = 64) go to the tensor-core kernel -- k <= 16 on its register lists, larger k through its
 * sampled + collect passes -- the rest to the integer-pipe kernel (LOP3 + POPC / DP4A).  On the tensor cores
 * both metrics are +-1 contractions in 4-bit floats (tcgen05.mma kind::mxf4: rows expanded bit -> e2m1 in shared
 * memory, fp32 sums of integers below 2^24: exact; HAMMING with unit block scales, XORSUM with the bit weights
 * 2^(7-p) as the query operand's block scales; XORSUM at W = 32 as weighted int8, kind::i8, exact int32).
 * 1 = integer-pipe kernel only; 2 / 3 = tensor-core kernel with one CTA / a CTA pair per query tile whenever
 * the call is eligible (any batch size); 5 / 6 = the same in the int8 form.  Results are bit-identical on every path (tests/test_gpu_mma_scan.py).  Replaces the same
 * reference lines as vl_search: vlite/model.py:71-74 + :84-85. */
int vl_set_scan_path(vl_corpus* c, int path);

/* ---- shard group: several shards searched as one corpus, in one process --- */
/* The in-process twin of the one-process-per-GPU path (the reference has no
 * multi-GPU counterpart).  A group borrows corpus handles -- usually one per GPU, any mix of
 * devices works -- whose base rows the caller has set (vl_corpus_set_base_row) so that global rows
 * are disjoint.  vl_shard_group_search replicates the queries to every shard, lets all shards scan
 * concurrently on their own streams -- each shard's last merge kernel stores its k candidates
 * straight into the first shard's device memory (peer stores over NVLink; a peer copy when the pair
 * has no peer access) -- and merges them there by (score, global row): the
 * result equals vl_search over the concatenated corpus.  Host pointers only; the call returns
 * when the results are in out_*.  filter_masks_or_null: NULL, or one mask pointer (or NULL) per
 * shard, each in that shard's local row numbering.  A group does not own its shards: destroy the
 * group first, then the handles.  One call at a time per group. */
typedef struct vl_shard_group vl_shard_group;
int vl_shard_group_create(int n_shards, vl_corpus* const* shards, vl_shard_group** out);
int vl_shard_group_destroy(vl_shard_group* g);
/* Optional, for shards that grow by interleaved appends: segment i of `shard` covers its local rows
 * from seg_local[i] (up to the next start) and global rows from seg_global[i]; both tables ascend
 * and the shard's base row must be 0.  Results then carry those global rows (insertion order), so
 * they do not depend on how batches were placed.  n_segments = 0 switches back to base rows. */
int vl_shard_group_set_segments(vl_shard_group* g, int shard, const uint64_t* seg_local,
                                const uint64_t* seg_global, int n_segments);
int vl_shard_group_rows(const vl_shard_group* g, uint64_t* rows, uint64_t* live_rows);
int vl_shard_group_search(vl_shard_group* g, const uint8_t* queries_host, int n_queries, int k, int metric,
                          const uint32_t* const* filter_masks_or_null, uint32_t* out_score_host,
                          uint64_t* out_row_host);

/* ---- exchange: the one exchange step of the row-sharded path for ONE PROCESS PER GPU -------- */
/* (the reference has no multi-GPU path, so this replaces nothing there.)  Every rank
 * stores its k candidates per query straight into every peer's gather buffer over NVLink (CUDA IPC
 * mappings opened once) and raises a flag; the merge kernel of each rank acquires all flags and merges by
 * (score, global row).  Same result as an all-gather + vl_merge_topk, without a collective library call
 * on the critical path.  Setup: every rank creates its exchange, publishes its 64-byte handle
 * (e.g. torch.distributed.all_gather_object) and attaches every peer's. */
typedef struct vl_exchange vl_exchange;
int vl_exchange_create(int device, int world, int rank, uint64_t slot_bytes, vl_exchange** out);
int vl_exchange_handle(vl_exchange* x, void* handle_out_64);
int vl_exchange_attach(vl_exchange* x, int peer_rank, const void* handle_64);
/* local_block_dev = [Q*k uint32 scores | Q*k uint64 GLOBAL rows] (Q*k even, 16 B aligned), block_bytes =
 * 12*Q*k <= slot_bytes.  Enqueues push + merge on cuda_stream; out_* get the merged top-k of all ranks. */
int vl_exchange_gather_merge(vl_exchange* x, void* cuda_stream, const void* local_block_dev,
                             uint64_t block_bytes, int n_queries, int k, uint32_t* out_score_dev,
                             uint64_t* out_row_dev);
int vl_exchange_destroy(vl_exchange* x);

#ifdef __cplusplus
}
#endif
#endif /* VLITE_B200_H */
