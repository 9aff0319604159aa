/*
 * smqtk_b200.h -- C ABI of the B200-native LSH nearest-neighbour hot path.
 *
 * One shared library (libsmqtk_b200.so), plain C linkage, raw device pointers
 * and sizes only (no torch / C++ types).  This is the boundary a maintainer of
 * SMQTK-Indexing binds (ctypes stub in INTEGRATION.md) to replace the pure
 * Python/numpy arithmetic of:
 *
 *   stage 1  ItqFunctor.get_hash             smqtk_indexing/impls/lsh_functor/itq.py:389-408
 *            + bit_vector_to_int_large       smqtk_indexing/utils/bits.py:4-20
 *   stage 2  LinearHashIndex._nn             smqtk_indexing/impls/hash_index/linear.py:206-244
 *            + hamming_distance              smqtk_indexing/utils/metrics.py:140-155
 *   stage 3  LSHNearestNeighborIndex._nn re-rank   smqtk_indexing/impls/nn_index/lsh.py:507-519
 *            + euclidean/cosine/hik          smqtk_indexing/utils/metrics.py:49-137
 *   build    LinearHashIndex._build_index / LSH _build_index (unique codes, code->rows)
 *                                            linear.py:148-165, lsh.py:316-329
 *   train    ItqFunctor.fit GEMMs            itq.py:338-383, 239-289
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *     the library never synchronises, allocates or frees caller-visible memory:
 *     scratch space is passed in, its size comes from the *_workspace_bytes query;
 *   - return value 0 = success, otherwise an SB_ERR_* code; sb_last_error()
 *     gives a thread-local message;
 *   - no global mutable state on the compute path: calls on different streams may run concurrently
 *     (process-wide bookkeeping only: the launch counter of sb_launch_count and the optional
 *     per-launch event records of sb_profile_enable / sb_profile_fetch).
 *
 * Hash-code layout (device): uint32[rows][W], each row is the INTEGER VALUE of
 * the reference's big-endian bit vector (bits.py:17-20) written in W 32-bit
 * words, word 0 most significant; W in {1,2,4,8,16,32}.  Bit j of a b-bit code
 * is integer bit p=b-1-j: word W-1-p/32, bit p%32.
 *
 * Packed result key: uint64 = (hamming_distance << 40) | global_row, so that
 * unsigned comparison is the canonical order "ascending distance, ties by row";
 * SB_KEY_EMPTY marks "no neighbour" (fewer than k rows).
 */
#ifndef SMQTK_B200_H
#define SMQTK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SB_API __attribute__((visibility("default")))
#else
#define SB_API
#endif

#define SB_OK 0
#define SB_ERR_INVALID_ARGUMENT 1
#define SB_ERR_CUDA 2
#define SB_ERR_WORKSPACE 3
#define SB_ERR_UNSUPPORTED 4

#define SB_KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
#define SB_KEY_ROW_BITS 40

/* norm_kind for sb_itq_hash (ItqFunctor(normalize=...), itq.py:172-191) */
#define SB_NORM_NONE 0  /* normalize=None                         */
#define SB_NORM_LP 1    /* numpy.linalg.norm(v, ord=p), p finite >0 */
#define SB_NORM_INF 2   /* ord=inf: max |v_i|                     */
#define SB_NORM_L0 3    /* ord=0: count of non-zeros              */

/* metric for sb_rerank / sb_pairwise_distance (lsh.py:236-255) */
#define SB_METRIC_EUCLIDEAN 0 /* metrics.py:73-86   sqrt(sum (a-b)^2)                 */
#define SB_METRIC_COSINE 1    /* metrics.py:120-137 2*acos(clip(cos_sim,-1,1))/pi      */
#define SB_METRIC_HIK 2       /* metrics.py:49-70   1 - sum min(a,b)                   */

SB_API int sb_version(void);
SB_API const char* sb_last_error(void);
/* Number of kernels this library has launched in the calling process (bench bookkeeping). */
SB_API uint64_t sb_launch_count(void);
/* Per-launch device timing for bench.py (no profiler needed): while enabled, every
 * kernel launch of this library is bracketed by CUDA events on its stream.
 * sb_profile_fetch waits for them, returns the number of launches recorded since
 * the last fetch and fills up to `cap` (kernel name, milliseconds) pairs in launch
 * order; `names[i]` points at static storage. */
SB_API int sb_profile_enable(int on);
SB_API int64_t sb_profile_fetch(const char** host_names, float* host_ms, int64_t cap);

/* ---- stage 1: ITQ hashing ------------------------------------------------
 * codes[r] = pack( ((X[r] / norm(X[r])) - mean) . R >= 0 )        itq.py:404-408
 * X: f32[n][ldx] (first D columns used); mean: f32[D] or NULL (= zeros);
 * R: f32[D][b] row-major; codes_out: u32[n][W] (W >= ceil(b/32), layout above);
 * z_out: optional f32[n][b] projections (tests / epsilon masks), may be NULL.
 * FP32 FFMA kernel, any shape.  variant: 0 or 1 (reserved).
 */
SB_API int sb_itq_hash(const float* X, int64_t n, int32_t D, int64_t ldx,
                const float* mean, const float* R, int32_t b,
                int32_t norm_kind, float norm_p,
                uint32_t* codes_out, int32_t W, float* z_out,
                int32_t variant, void* stream);

/* Tensor-core hashing (tcgen05.mma kind::tf32, 3xTF32 split, FP32 accumulation in
 * tensor memory, fused sign/bit-pack epilogue) for aligned shapes:
 *   D % 16 == 0, b % 32 == 0, 32 <= b <= 256, ldx % 4 == 0, 16-byte aligned pointers.
 * The rotation is pre-split ONCE per model into the kernel's shared-memory image
 * (hi/lo TF32 parts, UMMA core-matrix order): image_out holds
 * sb_itq_rotation_image_bytes(D, b) bytes (0 = shape not supported).
 * row_div: optional f32[n] row divisors (the norms of itq.py:184-189, 1 where 0),
 * NULL = no normalisation.  Same outputs as sb_itq_hash. */
SB_API int sb_itq_row_div(const float* X, int64_t n, int32_t D, int64_t ldx,
                   int32_t norm_kind, float norm_p, float* div_out, void* stream);
SB_API size_t sb_itq_rotation_image_bytes(int32_t D, int32_t b);
SB_API int sb_itq_rotation_image(const float* R, int32_t D, int32_t b, void* image_out, void* stream);
SB_API int sb_itq_hash_tc(const float* X, int64_t n, int32_t D, int64_t ldx,
                   const float* mean, const void* r_image, int32_t b, const float* row_div,
                   uint32_t* codes_out, int32_t W, float* z_out, void* stream);

/* ---- stage 2: Hamming scan + top-k -----------------------------------------
 * For each of Q query codes: the k rows of db with the smallest
 * popcount(q xor row), canonical order (distance, idx_base + row).
 * db: u32[U][W]; q: u32[Q][W]; keys_out: u64[Q][k] ascending, padded with
 * SB_KEY_EMPTY when U < k.   linear.py:232-240 (heapq.nsmallest over the set).
 */
SB_API size_t sb_hamming_scan_workspace_bytes(int64_t U, int32_t W, int32_t Q, int32_t k);
SB_API int sb_hamming_scan(const uint32_t* db, int64_t U, int32_t W,
                    const uint32_t* q, int32_t Q, int32_t k, int64_t idx_base,
                    uint64_t* keys_out, void* workspace, size_t workspace_bytes,
                    void* stream);
/* Same, choosing the popcount formulation: variant 0 = default (carry-save adder
 * tree + 5 POPC per 256 bits), 1 = plain POPC per word.  Results are identical. */
SB_API int sb_hamming_scan_variant(const uint32_t* db, int64_t U, int32_t W,
                            const uint32_t* q, int32_t Q, int32_t k, int64_t idx_base,
                            uint64_t* keys_out, void* workspace, size_t workspace_bytes,
                            int32_t variant, void* stream);
/* sb_hamming_scan, predicated on a DEVICE flag: every kernel returns at once unless *enable != 0,
 * in which case keys_out is overwritten with the exact XOR/POPC result.  Launched right after
 * sb_hamming_scan_tc with its overflow flag it makes the tensor-core scan exact for any table
 * without a host round trip (workspace: sb_hamming_scan_workspace_bytes). */
SB_API int sb_hamming_scan_if(const uint32_t* db, int64_t U, int32_t W, const uint32_t* q, int32_t Q, int32_t k,
                       int64_t idx_base, uint64_t* keys_out, const int32_t* enable, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Batched scan on the tensor cores (hamming_tc.cu): bits -> +-1 FP8 E4M3, tcgen05.mma
 * kind::f8f6f4, dot = 32 W - 2 * distance -- exact integers in the FP32 accumulator.  Same
 * contract and the same keys as sb_hamming_scan, for W in {1, 2, 4, 8} and k <= 256; meant for
 * large query batches, where the XOR/POPC scan is bound by the integer pipes instead of HBM.
 * *overflow_out (device int32) becomes 1 if a per-query candidate buffer overflowed: the keys
 * must then be discarded and sb_hamming_scan run instead. */
SB_API int sb_hamming_scan_tc_supported(int64_t U, int32_t W, int32_t Q, int32_t k);
SB_API size_t sb_hamming_scan_tc_workspace_bytes(int64_t U, int32_t W, int32_t Q, int32_t k);
SB_API int sb_hamming_scan_tc(const uint32_t* db, int64_t U, int32_t W,
                       const uint32_t* q, int32_t Q, int32_t k, int64_t idx_base,
                       uint64_t* keys_out, int32_t* overflow_out,
                       void* workspace, size_t workspace_bytes, void* stream);
/* The same scan with PACKED FP4 operands (hamming_tc4.cu; the one the host side calls by default): bits -> +-1
 * E2M1, tcgen05.mma kind::mxf4.block_scale, K = 64 per MMA at the cycle count of kind::f8f6f4's K = 32, three
 * queries per FP32 accumulator column (8-bit fields; the b / c queries' products scaled by 2^8 / 2^16 through the
 * block scales).  The same exact integer distances, the same keys, the same overflow flag (also raised when a
 * threshold above 252 leaves no room for the field windows).  Same contract as sb_hamming_scan_tc; k <= 256. */
SB_API int sb_hamming_scan_tc4_supported(int64_t U, int32_t W, int32_t Q, int32_t k);
SB_API size_t sb_hamming_scan_tc4_workspace_bytes(int64_t U, int32_t W, int32_t Q, int32_t k);
SB_API int sb_hamming_scan_tc4(const uint32_t* db, int64_t U, int32_t W, const uint32_t* q, int32_t Q, int32_t k,
                        int64_t idx_base, uint64_t* keys_out, int32_t* overflow_out, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Merge `parts` sorted key lists per query (keys_in: u64[parts][Q][k], e.g. the
 * all-gathered per-GPU results) into the global top-k and decode:
 * out_dist i32[Q][k] (-1 = empty), out_idx i64[Q][k] (-1 = empty); either of
 * out_keys (u64[Q][k]) / out_dist / out_idx may be NULL. */
SB_API int sb_topk_merge(const uint64_t* keys_in, int32_t parts, int32_t Q, int32_t k,
                  uint64_t* out_keys, int32_t* out_dist, int64_t* out_idx, void* stream);

/* Convenience: sb_hamming_scan + decode in one call (single device). */
SB_API int sb_hamming_topk(const uint32_t* db, int64_t U, int32_t W,
                    const uint32_t* q, int32_t Q, int32_t k, int64_t idx_base,
                    int32_t* out_dist, int64_t* out_idx,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- stage 3: candidate re-rank ----------------------------------------------
 * out[j] = metric(q[query_of(j)], db[cand_idx[j]]) for j in [cand_off[0], cand_off[Q]),
 * candidates of query i are j in [cand_off[i], cand_off[i+1]).  Element math is
 * FP32 FFMA with error-free (hi, lo) accumulation, the per-candidate scalar tail
 * (sqrt / divide / acos) is FP64; lsh.py:507-511 + metrics.py.
 * db: f32[N][ldd]; q: f32[Q][ldq]; cand_idx: i64[M]; cand_off: i64[Q+1]; out: f64[M]
 * (NaN for an out-of-range row, and for cosine against a zero vector as in the reference).
 */
SB_API int sb_rerank(const float* db, int64_t N, int32_t D, int64_t ldd,
              const float* q, int32_t Q, int64_t ldq,
              const int64_t* cand_idx, const int64_t* cand_off, int64_t M,
              int32_t metric, double* out, void* stream);

/* Like sb_rerank with cand_idx holding GLOBAL rows of a table whose rows [row_base, row_base + N)
 * are in db (NaN outside). */
SB_API int sb_rerank_base(const float* db, int64_t N, int64_t row_base, int32_t D, int64_t ldd,
                   const float* q, int32_t Q, int64_t ldq,
                   const int64_t* cand_idx, const int64_t* cand_off, int64_t M,
                   int32_t metric, double* out, void* stream);

/* Row-sharded variant: db holds the GLOBAL rows [row_base, row_base + N); candidates
 * owned by another shard (and -1 padding) get 0.0, so that one all-reduce(SUM) over
 * the shards assembles every distance exactly (x + 0 + ... + 0 == x). */
SB_API int sb_rerank_shard(const float* db, int64_t N, int64_t row_base, int32_t D, int64_t ldd,
                    const float* q, int32_t Q, int64_t ldq,
                    const int64_t* cand_idx, const int64_t* cand_off, int64_t M,
                    int32_t metric, double* out, void* stream);

/* Per query: order its candidates by (distance, candidate position) and keep
 * the first n (lsh.py:513-519: stable sort by distance, slice).  out_pos i64[Q][n]
 * = position j into cand_idx (-1 = fewer than n candidates), out_dist f64[Q][n]
 * (NaN distances order last). */
SB_API int sb_rerank_select(const double* dist, const int64_t* cand_off, int32_t Q, int32_t n,
                     int64_t* out_pos, double* out_dist, void* stream);

/* Same selection, fused with the row gather: out_rows[q][i] = cand_idx[position]
 * (-1 = fewer than n candidates).  cand_cnt (optional, i64[Q]) limits each query to
 * the first cand_cnt[q] entries of its segment (fixed-pitch layout below).
 * tie_by_row != 0: equal distances are ordered by candidate row instead of position. */
SB_API int sb_rerank_select_rows(const double* dist, const int64_t* cand_off, const int64_t* cand_cnt,
                          const int64_t* cand_idx, int32_t Q, int32_t n, int32_t tie_by_row,
                          int64_t* out_rows, double* out_dist, void* stream);

/* Re-rank against a ROW-SHARDED descriptor table spread over the GPUs of one box (BASELINE north_star:
 * "database sharded by row across the 8 GPUs"): shards f32*[n_shards] is a DEVICE array of pointers valid
 * in the calling process -- the local shard plus CUDA-IPC mappings of the peers' shards --, shard s holds
 * global rows [shard_bounds[s], shard_bounds[s+1]) (i64[n_shards+1], device), all with leading dimension
 * ldd.  Each candidate row is read where it lives (NVLink loads for peer shards): the distances of
 * lsh.py:507-512 without any collective.  aligned16 != 0 promises 16-byte aligned shard bases.
 * sb_enable_peer_access(peer): cudaDeviceEnablePeerAccess from the current device (idempotent). */
SB_API int sb_rerank_peer(const float* const* shards, const int64_t* shard_bounds, int32_t n_shards, int32_t D,
                   int64_t ldd, const float* q, int32_t Q, int64_t ldq, const int64_t* cand_idx,
                   const int64_t* cand_off, int64_t M, int64_t pitch, int32_t metric, int32_t aligned16, double* out,
                   void* stream);
/* sb_rerank for the fixed-pitch candidate layout of sb_expand_candidates (M = Q * pitch; the query of slot j
 * is j / pitch, no search); `pitch` in sb_rerank_peer means the same (0 = ragged lists, cand_off is searched). */
SB_API int sb_rerank_pitched(const float* db, int64_t N, int32_t D, int64_t ldd, const float* q, int32_t Q, int64_t ldq,
                      const int64_t* cand_idx, const int64_t* cand_off, int64_t pitch, int32_t metric, double* out,
                      void* stream);
SB_API int sb_enable_peer_access(int32_t peer_device);
/* CUDA IPC plumbing for the shard table of sb_rerank_peer (one process per GPU).  sb_ipc_export: the 64-byte
 * handle (host memory) of the cudaMalloc allocation that contains dev_ptr + dev_ptr's offset in it.
 * sb_ipc_import: opens a peer's handle under the CURRENT device with lazy peer access, so that this device's
 * kernels can load from it; returns the allocation base in this process.  sb_ipc_release closes it. */
SB_API int sb_ipc_export(const void* dev_ptr, void* handle_out, int64_t* offset_out);
SB_API int sb_ipc_import(const void* handle, void** base_out);
SB_API int sb_ipc_release(void* base);

/* Candidate expansion (lsh.py:490-496): the descriptor rows of each query's near codes,
 * in (code rank, row) order.  code_rows i64[Q][n] = rows of the unique-code table
 * (-1 = none); csr_off i64[U+1] / csr_rows i64[N] = code -> rows map.  Output in a
 * FIXED-PITCH layout (no host round trip for the total): cand_idx i64[Q][pitch] padded
 * with -1, cand_off i64[Q+1] = q*pitch, cand_cnt i64[Q] = rows present.  pitch must be
 * >= n * (max rows per code); surplus rows are dropped.  n <= 2048. */
SB_API int sb_expand_candidates(const int64_t* code_rows, int32_t Q, int32_t n,
                         const int64_t* csr_off, const int64_t* csr_rows, int64_t pitch,
                         int64_t* cand_idx, int64_t* cand_off, int64_t* cand_cnt, void* stream);

/* ---- index build: sorted-unique code table + code -> rows CSR (replaces the Python set of ints of
 * LinearHashIndex._build_index / _update_index / _remove_from_index, linear.py:148-204, and the
 * {hash int: set(uuid)} loop of LSHNearestNeighborIndex._build_index, lsh.py:316-329).
 *   codes u32[n_rows][W] per-row codes (layout above); rows_in i64[n] = the rows to index (live rows,
 *   ascending) or NULL for all n = n_rows rows.
 * Outputs (capacity n unless noted): table_out u32[U][W] ascending by integer value, distinct;
 *   row_code_out i64[n_rows] row -> table row (-1 for rows not in rows_in); csr_off_out i64[U+1]
 *   (capacity n+1), csr_rows_out i64[n]: rows of code c are csr_rows[csr_off[c] .. csr_off[c+1]) in
 *   ascending row order; stats_out i64[2] = {U, most rows sharing one code} (device memory: the
 *   caller reads it once to learn U).
 * One stable LSD radix sort of a row permutation (8-bit digits, one kernel per pass with decoupled
 * look-back; digits all rows agree on are skipped) + boundary flags + scan + scatter; no library
 * sort.  n < 2^30, n_rows < 2^32, W in {1,2,4,8,16,32}. */
SB_API size_t sb_unique_codes_workspace_bytes(int64_t n, int32_t W);
SB_API int sb_unique_codes(const uint32_t* codes, int64_t n_rows, int32_t W, const int64_t* rows_in, int64_t n,
                    uint32_t* table_out, int64_t* row_code_out, int64_t* csr_off_out, int64_t* csr_rows_out,
                    int64_t* stats_out, void* workspace, size_t workspace_bytes, void* stream);

/* Selection over LARGE candidate lists (reference: the O(m log m) `sorted(...)` of lsh.py:513-516):
 * same contract as sb_rerank_select_rows (cand_cnt and cand_idx optional; cand_idx == NULL writes
 * candidate positions), by ONE stable radix sort of (query, distance[, row]) keys over all M
 * candidates instead of a per-query rank count.  NaN distances sort last.  tie_by_row needs
 * candidate rows < 2^32.  M < 2^30. */
SB_API size_t sb_rerank_select_sorted_workspace_bytes(int64_t M);
SB_API int sb_rerank_select_sorted(const double* dist, const int64_t* cand_off, const int64_t* cand_cnt,
                            const int64_t* cand_idx, int64_t M, int32_t Q, int32_t n, int32_t tie_by_row,
                            int64_t* out_pos, double* out_dist, void* workspace, size_t workspace_bytes, void* stream);

/* Exhaustive Hamming top-k for ANY k (the reference's heapq.nsmallest takes any n, linear.py:232-240;
 * the scan kernels keep lists of <= 2048): every (query, row) key is materialised and radix-sorted.
 * keys_out u64[Q][k] packed keys ascending, SB_KEY_EMPTY padded when k > U.  Q * U < 2^30 per call
 * (the caller loops over query chunks).  W in {1,2,4,8,16,32}. */
SB_API size_t sb_hamming_topk_sorted_workspace_bytes(int64_t U, int32_t Q);
SB_API int sb_hamming_topk_sorted(const uint32_t* db, int64_t U, int32_t W, const uint32_t* q, int32_t Q, int32_t k,
                           int64_t idx_base, uint64_t* keys_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- flat index: exact L2 k-nearest rows (reference impls/nn_index/faiss.py:751-831
 * with 'IDMap,Flat' + metrics.py:73-86).  Tensor-core filter (3xTF32 |x|^2+|q|^2-2x.q against
 * per-query thresholds, chunked) + exact FP32 error-free re-rank of the survivors; result =
 * the k rows with the smallest euclidean distance in (distance, row) order.
 *   sb_l2_prepare: xn_out f32[N] = |row|^2, xn_max_out f32[1] = max (once per table).
 *   sb_l2_topk:    out_idx i64[Q][k] (-1 = fewer than k rows), out_dist f64[Q][k] (NaN);
 *                  overflow_out i32[Q] != 0 marks a query whose survivor buffer overflowed
 *                  (massive near-ties): the caller re-does it with sb_rerank over all rows.
 * Needs D % 16 == 0, ldd % 4 == 0, 16-byte aligned db, k <= 256 (sb_l2_topk_supported). */
SB_API int sb_l2_prepare(const float* db, int64_t N, int32_t D, int64_t ldd,
                  float* xn_out, float* xn_max_out, void* stream);
SB_API int sb_l2_topk_supported(int64_t N, int32_t D, int64_t ldd, int32_t k);
SB_API size_t sb_l2_topk_workspace_bytes(int32_t D, int32_t Q, int32_t k);
SB_API int sb_l2_topk(const float* db, int64_t N, int32_t D, int64_t ldd,
               const float* xn, const float* xn_max,
               const float* q, int32_t Q, int64_t ldq, int32_t k,
               int64_t* out_idx, double* out_dist, int32_t* overflow_out,
               void* workspace, size_t workspace_bytes, void* stream);

/* ---- training: N-scaled contractions of ItqFunctor.fit (itq.py:338-383, 239-289) ----
 * FP64 accumulation, deterministic two-stage row reductions.  Matrix operands are
 * (pointer, kind, row pitch, columns, div, mean[, bits]):
 *   kind 0 = f32, kind 1 = f64: element value = x / div[r] - mean[c] (div, mean optional: NULL)
 *   kind 2 = packed sign bits u32[n][pitch] of `bits`-bit codes: value = bit ? +1 : -1
 */
/* div_out[r] = norm(X[r]) (1 where the norm is 0); itq.py:184-189 */
SB_API int sb_fit_row_div(const void* X, int32_t x_kind, int64_t n, int32_t D, int64_t ldx,
                          int32_t norm_kind, double norm_p, double* div_out, void* stream);
/* scratch size for sb_fit_col_mean / sb_fit_gram with output Ma x Mb over n rows */
SB_API size_t sb_fit_workspace_bytes(int64_t n, int32_t Ma, int32_t Mb);
/* mean_out[c] = mean_r (X[r][c] / div[r]); itq.py:343 */
SB_API int sb_fit_col_mean(const void* X, int32_t x_kind, int64_t n, int32_t D, int64_t ldx,
                           const double* div, double* mean_out,
                           void* workspace, size_t workspace_bytes, void* stream);
/* out f64[Ma][Mb] = scale * sum_r opA(r)^T opB(r): covariance (itq.py:351) and UX^T V (itq.py:274) */
SB_API int sb_fit_gram(const void* A, int32_t a_kind, int64_t lda, int32_t Ma,
                       const double* a_div, const double* a_mean, int32_t a_bits,
                       const void* B, int32_t b_kind, int64_t ldb, int32_t Mb,
                       const double* b_div, const double* b_mean, int32_t b_bits,
                       int64_t n, double scale, double* out,
                       void* workspace, size_t workspace_bytes, void* stream);
/* opA[n][K] . Bm f64[K][M] -> out_f64 f64[n][M] and/or sign bits out_codes u32[n][Wc]
 * (projection itq.py:378; Z = V.R with the sign step itq.py:270-272, 285-287) */
SB_API int sb_fit_project(const void* A, int32_t a_kind, int64_t lda, int32_t K,
                          const double* a_div, const double* a_mean,
                          const double* Bm, int32_t M, int64_t n,
                          double* out_f64, uint32_t* out_codes, int32_t Wc, void* stream);

/* Tensor-core form of the ITQ iteration's row-scaled product (itq.py:274 in the V-free form of
 * fit.py): out f64[b][D] = sum_r (bit_m(codes[r]) ? +1 : -1) * (X[r][d] / row_div[r] - mean[d]).
 * codes u32[n][W] (layout above), X f32[n][ldx]; mean f32[D] / row_div f32[n] optional.
 * `bound` >= max |X / row_div - mean| (any finite upper bound; a tight one keeps more bits): the
 * centred values are split on fixed-point grids (2^-10 and 2^-21 of the bound's power of two), so
 * the TF32 products and the FP32 accumulation in tensor memory are exact; windows of 8192 rows are
 * flushed to FP64 and reduced in a fixed order (deterministic).  Error: the 2^-22 * bound rounding
 * of each centred value, nothing else.  Needs D % 32 == 0, ldx % 4 == 0, b % 4 == 0, b <= 256. */
SB_API int sb_fit_gram_bits_tc_supported(int64_t n, int32_t D, int64_t ldx, int32_t b);
SB_API size_t sb_fit_gram_bits_tc_workspace_bytes(int64_t n, int32_t D, int32_t b);
SB_API int sb_fit_gram_bits_tc(const uint32_t* codes, int32_t W, int32_t b, const float* X, int64_t n, int32_t D,
                        int64_t ldx, const float* mean, const float* row_div, float bound, double* out,
                        void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMQTK_B200_H */
