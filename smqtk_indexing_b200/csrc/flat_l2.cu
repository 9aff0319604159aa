// Exact flat (brute-force) L2 k-nearest-neighbour search -- SURVEY.md section 8(f) row N1,
// BASELINE.json config 3.  Semantics of the reference's flat index
// (smqtk_indexing/impls/nn_index/faiss.py:751-831 with factory 'IDMap,Flat', metric L2:
// the k rows with the smallest squared-L2 distance, reported as sqrt distances and
// re-ordered by smqtk_indexing/utils/metrics.py:73-86 euclidean_distance), made exact:
// the oracle is a float64 brute force with ties broken by row.
//
// Two stages, both on the device, no host round trip:
//   1. FILTER on the tensor cores (itq_hash_tc.cu, EPI_L2): approximate d2 = |x|^2 + |q|^2 -
//      2 x.q with a single-pass TF32 GEMM (256 rows x 256 queries per tile; the threshold
//      terms ride in one extra, fully split K chunk so the test is a sign test), compared
//      against a per-query threshold tq = tau + margin.  The database is walked in chunks that grow
//      8x (2048, 16k, 131k, 1M, 8.4M, ... rows); after each chunk a per-query bitonic sort
//      of the survivors tightens tau to the k-th smallest approximate distance seen so far,
//      so a chunk appends only ~8k survivors per query (the first chunk keeps every row).
//      margin = eps * (|q|^2 + max|x|^2) covers twice the approximation error, hence every
//      true top-k row survives every threshold (tau is an upper bound of the approximate
//      k-th distance of a SUBSET of the rows, which is >= the true k-th distance - error).
//   2. EXACT: the survivors (a few hundred per query) are re-ranked with the direct-form
//      error-free FP32 distance kernel (rerank.cu) and the k best are selected by
//      (distance, row).
// A query whose buffer overflows (massive near-ties) is flagged; the host wrapper re-does
// those queries with the exact kernel over every row.
#include <math.h>

#include "common.cuh"

namespace sb {
int tc_l2_tma_supported(int32_t D, int64_t ldd, const float* db);
size_t tc_l2_tma_image_bytes(int32_t D, int col_blocks);
int tc_l2_tma_query_image(const float* q, int Q, int D, long long ldq, int col_blocks, void* img, float* qn, cudaStream_t st);
int tc_l2_tma_threshold_image(int Q, int cols, int D, const float* qn, const float* tq, void* img, cudaStream_t st);
int tc_l2_filter_tma(const float* X, int64_t n, int32_t D, int64_t ldx, const void* image, int col_blocks, const float* xn,
                     const float* tq, unsigned long long* cand_buf, int* cand_cnt, int cap, unsigned row_base,
                     int dense, cudaStream_t st);
int tc_l2_filter(const float* X, int64_t n, int32_t D, int64_t ldx, const uint32_t* q_image, int col_blocks,
                 const float* xn, const float* tq, unsigned long long* cand_buf, int* cand_cnt, int cap,
                 unsigned row_base, int passes, cudaStream_t st);
}

namespace {

constexpr int KC = 16, KCHUNKS = 4;       // must match itq_hash_tc.cu
constexpr int QB = 256;                   // query columns per block
// margin = eps * (|q|^2 + max|x|^2) must cover TWICE the error of the approximate d2.
//   3xTF32: measured |dz| <= 1e-6 |x||q| (tests/test_gpu_kernels.py HASH_EPS = 1e-5 is the asserted bound)
//           -> d2 error <= 2e-5 |x||q| <= 1e-5 (|x|^2 + |q|^2)            -> eps = 4e-5 (2x slack)
//   1xTF32: operands rounded to 11 bits: |dz| <= (2^-11 + 2^-11 + 2^-22) sum|x_i q_i| <= 2^-10 |x||q|
//           (Cauchy-Schwarz) -> d2 error <= 2^-9 |x||q| <= 2^-10 (|x|^2 + |q|^2) -> eps = 2^-9 + 4e-5
constexpr float L2_EPS_3X = 4e-5f;
constexpr float L2_EPS_1X = 0.001953125f + 4e-5f;
//   TMA variant: A is raw FP32 TRUNCATED to TF32 by the tensor core (2^-10), B rounded (2^-11):
//           |dz| <= 1.5 * 2^-10 |x||q| -> d2 error <= 1.5 * 2^-10 (|x|^2 + |q|^2)   -> eps = 3 * 2^-10 + 4e-5
constexpr float L2_EPS_TMA = 0.0029296875f + 4e-5f;
constexpr int L2_PASSES = 1;              // single TF32 pass on the data chunks (the exact stage restores exactness)
constexpr int FIRST_CHUNK_MAX = 2048;
constexpr int GROWTH = 8;

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// |x|^2 per row (one warp per row) and the maximum over rows.
__global__ void __launch_bounds__(256)
row_sqnorm_kernel(const float* __restrict__ X, long long n, int D, long long ldx, float* __restrict__ xn,
                  int* __restrict__ xn_max_bits) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* x = X + row * ldx;
  double acc = 0.0;
  for (int d = lane; d < D; d += 32) acc += (double)x[d] * (double)x[d];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(sb::FULL_MASK, acc, o);
  if (lane == 0) {
    const float v = (float)acc;
    xn[row] = v;
    atomicMax(xn_max_bits, __float_as_int(v));    // non-negative floats order like ints
  }
}

// Queries -> per-block hi/lo TF32 images in the UMMA core-matrix order (layout of
// rotation_image_kernel with R = Q^T, b = 256; columns past Q are zero) and |q|^2.
__global__ void query_image_kernel(const float* __restrict__ q, int Q, int D, long long ldq, int col_blocks,
                                   uint32_t* __restrict__ img, float* __restrict__ qn) {
  const long long total = (long long)col_blocks * QB * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % D);
    const long long col = i / D;                  // global query column
    const int jb = (int)(col / QB), n = (int)(col % QB);
    const float r = (col < Q) ? q[col * ldq + k] : 0.0f;
    const uint32_t hi = tf32_rna(r);
    const uint32_t lo = tf32_rna(r - __uint_as_float(hi));
    const int kc = k / KC, c = (k % KC) / 4, e = k & 3;
    const size_t base = (size_t)jb * (D + KC) * QB * 2 + ((size_t)kc * 2) * KCHUNKS * QB * 4;
    const size_t off = ((size_t)c * QB + n) * 4 + e;
    img[base + off] = hi;
    img[base + (size_t)KCHUNKS * QB * 4 + off] = lo;
  }
  // |q|^2: one thread per column (D is small)
  for (long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x; col < (long long)col_blocks * QB;
       col += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    if (col < Q)
      for (int k = 0; k < D; ++k) acc += (double)q[col * ldq + k] * (double)q[col * ldq + k];
    qn[col] = (float)acc;
  }
}

__global__ void l2_init_kernel(int Q, int cols, const float* __restrict__ qn, const int* __restrict__ xn_max_bits,
                               float* __restrict__ tau, float* __restrict__ margin, float* __restrict__ tq,
                               int* __restrict__ cnt, int* __restrict__ overflow, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cols) return;
  const float xmax = __int_as_float(*xn_max_bits);
  // no row can be farther than |q| + max|x|: a finite threshold that the whole first chunk passes
  const float far = sqrtf(qn[i]) + sqrtf(xmax);
  margin[i] = eps * (qn[i] + xmax);
  tau[i] = far * far * 1.0001f;
  tq[i] = tau[i] + margin[i];
  cnt[i] = 0;
  if (i < Q) overflow[i] = 0;
}

// The synthetic K chunk of every query block: B row = (1, -h, 0, ...) with h = (|q|^2 - tq)/2;
// padding columns get a huge h so that they never pass.  Rewritten before every filter pass.
__global__ void l2_threshold_image_kernel(int Q, int cols, int D, const float* __restrict__ qn,
                                          const float* __restrict__ tq, uint32_t* __restrict__ img) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= cols) return;
  const int jb = col / QB, n = col % QB;
  const float mh = (col < Q) ? -0.5f * (qn[col] - tq[col]) : -1.0e30f;
  const uint32_t one = __float_as_uint(1.0f);
  const uint32_t hi = tf32_rna(mh);
  const uint32_t lo = tf32_rna(mh - __uint_as_float(hi));
  const size_t base = (size_t)jb * (D + KC) * QB * 2 + ((size_t)(D / KC) * 2) * KCHUNKS * QB * 4;
  const size_t lo_part = (size_t)KCHUNKS * QB * 4;
  for (int c = 0; c < KCHUNKS; ++c)
    for (int e = 0; e < 4; ++e) {
      const size_t off = ((size_t)c * QB + n) * 4 + e;
      uint32_t vh = 0u, vl = 0u;
      if (c == 0 && e == 0) vh = one;
      if (c == 0 && e == 1) { vh = hi; vl = lo; }
      img[base + off] = vh;
      img[base + lo_part + off] = vl;
    }
}

// The first chunk keeps every row (no threshold yet), so it does not go through the filter:
// one CTA per query writes the keys of rows [0, m0) directly (direct-form FP32 d2, no atomics).
__global__ void __launch_bounds__(256)
l2_seed_kernel(const float* __restrict__ db, int m0, int D, long long ldd, const float* __restrict__ q, long long ldq,
               unsigned long long* __restrict__ buf, int* __restrict__ cnt, int cap) {
  extern __shared__ float s_q[];
  const int qi = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) s_q[d] = q[(long long)qi * ldq + d];
  __syncthreads();
  for (int r = threadIdx.x; r < m0; r += blockDim.x) {
    const float* x = db + (long long)r * ldd;
    float acc = 0.0f;
    for (int d = 0; d < D; ++d) {
      const float t = x[d] - s_q[d];
      acc = fmaf(t, t, acc);
    }
    buf[(long long)qi * cap + r] = ((unsigned long long)__float_as_uint(acc) << 32) | (unsigned long long)r;
  }
  if (threadIdx.x == 0) cnt[qi] = m0;
}

__global__ void l2_set_count_kernel(int* __restrict__ cnt, int Q, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Q) cnt[i] = v;
}

// One CTA per query: sort the survivors by (approximate d2, row), tighten tau to the k-th
// smallest, keep everything within the margin of it.  `final` also emits the candidate
// rows in the layout sb_rerank / sb_rerank_select_rows consume.
constexpr int CP_THREADS = 1024;

__global__ void __launch_bounds__(CP_THREADS)
l2_compact_kernel(unsigned long long* __restrict__ buf, int* __restrict__ cnt, int cap, int k, float* __restrict__ tau,
                  const float* __restrict__ margin, float* __restrict__ tq, int* __restrict__ overflow, int final,
                  long long* __restrict__ cand_idx, long long* __restrict__ cand_off, long long* __restrict__ cand_cnt,
                  int Q, int final_pitch) {
  extern __shared__ unsigned long long s_key[];   // P = next pow2 >= min(count, cap)
  const int qi = blockIdx.x, tid = threadIdx.x;
  const int raw = cnt[qi];
  const int m = min(raw, cap);
  if (raw > cap && tid == 0) overflow[qi] = 1;
  unsigned long long* mine = buf + (size_t)qi * cap;
  int P = 1;
  while (P < m) P <<= 1;
  for (int i = tid; i < P; i += CP_THREADS) s_key[i] = (i < m) ? mine[i] : ~0ull;
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < P / 2; i += CP_THREADS) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long a = s_key[lo], b = s_key[hi];
        if ((a > b) == up) { s_key[lo] = b; s_key[hi] = a; }
      }
      __syncthreads();
    }
  }
  float t = tau[qi];                               // fewer than k rows seen: the bound stands
  int keep = m;
  if (m >= k) {
    t = __uint_as_float((unsigned)(s_key[k - 1] >> 32));
    const float lim = t + margin[qi];
    // keys are sorted: binary search the first key whose d2 exceeds the limit
    int lo = k, hi = m;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__uint_as_float((unsigned)(s_key[mid] >> 32)) <= lim) lo = mid + 1; else hi = mid;
    }
    keep = lo;
  }
  for (int i = tid; i < keep; i += CP_THREADS) mine[i] = s_key[i];
  if (tid == 0) {
    cnt[qi] = keep;
    tau[qi] = t;
    tq[qi] = t + margin[qi];
  }
  if (final) {
    // the exact stage works on a tighter pitch than the survivor buffers: ~k + margin band rows remain
    long long* out = cand_idx + (size_t)qi * final_pitch;
    for (int i = tid; i < final_pitch; i += CP_THREADS) out[i] = (i < keep) ? (long long)(s_key[i] & 0xffffffffull) : -1ll;
    if (tid == 0) {
      if (keep > final_pitch) overflow[qi] = 1;
      cand_cnt[qi] = min(keep, final_pitch);
      cand_off[qi] = (long long)qi * final_pitch;
      if (qi == Q - 1) cand_off[Q] = (long long)Q * final_pitch;
    }
  }
}

struct L2Plan {
  int col_blocks, cols, cap, final_pitch;
  size_t off_img, off_qn, off_tau, off_margin, off_tq, off_cnt, off_buf, off_cand_idx, off_cand_off, off_cand_cnt,
      off_dist, total;
};

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

// Survivor slots per query.  A chunk GROWTH x the rows seen so far appends ~GROWTH * k survivors
// plus the rows inside the margin band (about as many again with the single-pass TF32 margins),
// on top of the ~k kept ones; 4 * (GROWTH + 1) * k leaves a 2x cushion over that.
int l2_cap(int k) {
  int cap = 2048;
  while (cap < 4 * (GROWTH + 1) * k) cap <<= 1;
  return cap;
}

L2Plan make_l2_plan(int32_t D, int32_t Q, int32_t k) {
  L2Plan p;
  p.col_blocks = (Q + QB - 1) / QB;
  p.cols = p.col_blocks * QB;
  p.cap = l2_cap(k);
  p.final_pitch = 512;
  while (p.final_pitch < 8 * k) p.final_pitch <<= 1;
  if (p.final_pitch > p.cap) p.final_pitch = p.cap;
  size_t o = 0;
  p.off_img = o;      o += align256((size_t)p.cols * (D + KC) * 2 * sizeof(uint32_t));
  p.off_qn = o;       o += align256((size_t)p.cols * sizeof(float));
  p.off_tau = o;      o += align256((size_t)p.cols * sizeof(float));
  p.off_margin = o;   o += align256((size_t)p.cols * sizeof(float));
  p.off_tq = o;       o += align256((size_t)p.cols * sizeof(float));
  p.off_cnt = o;      o += align256((size_t)p.cols * sizeof(int));
  p.off_buf = o;      o += align256((size_t)p.cols * p.cap * sizeof(unsigned long long));
  p.off_cand_idx = o; o += align256((size_t)Q * p.final_pitch * sizeof(long long));
  p.off_cand_off = o; o += align256((size_t)(Q + 1) * sizeof(long long));
  p.off_cand_cnt = o; o += align256((size_t)Q * sizeof(long long));
  p.off_dist = o;     o += align256((size_t)Q * p.final_pitch * sizeof(double));
  p.total = o;
  return p;
}

}  // namespace

extern "C" {

int sb_rerank_base(const float* db, int64_t N, int64_t row_base, int32_t D, int64_t ldd, const float* q, int32_t Q,
                   int64_t ldq, const int64_t* cand_idx, const int64_t* cand_off, int64_t M, int32_t metric,
                   double* out, void* stream);
int sb_rerank_select_rows(const double* dist, const int64_t* cand_off, const int64_t* cand_cnt, const int64_t* cand_idx,
                          int32_t Q, int32_t n, int32_t tie_by_row, int64_t* out_rows, double* out_dist, void* stream);

int sb_l2_prepare(const float* db, int64_t N, int32_t D, int64_t ldd, float* xn_out, float* xn_max_out, void* stream) {
  SB_REQUIRE(N >= 0 && D >= 1 && ldd >= D, "sb_l2_prepare: bad sizes");
  SB_REQUIRE(xn_max_out != nullptr && (N == 0 || (db != nullptr && xn_out != nullptr)), "sb_l2_prepare: NULL pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SB_CUDA_TRY(cudaMemsetAsync(xn_max_out, 0, sizeof(float), st));
  if (N == 0) return SB_OK;
  sb::ProfScope prof("row_sqnorm_kernel", st);
  row_sqnorm_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(db, N, D, ldd, xn_out, reinterpret_cast<int*>(xn_max_out));
  sb::count_launch();
  return sb::check_launch("row_sqnorm_kernel");
}

int sb_l2_topk_supported(int64_t N, int32_t D, int64_t ldd, int32_t k) {
  return N >= 1 && D >= KC && D % KC == 0 && ldd % 4 == 0 && k >= 1 && k <= 256 && N < (1ll << 32);
}

size_t sb_l2_topk_workspace_bytes(int32_t D, int32_t Q, int32_t k) {
  if (D < 1 || Q < 1 || k < 1) return 0;
  return make_l2_plan(D, Q, k).total;
}

int sb_l2_topk(const float* db, int64_t N, int32_t D, int64_t ldd, const float* xn, const float* xn_max,
               const float* q, int32_t Q, int64_t ldq, int32_t k,
               int64_t* out_idx, double* out_dist, int32_t* overflow_out, void* workspace, size_t workspace_bytes,
               void* stream) {
  SB_REQUIRE(Q >= 1 && ldq >= D, "sb_l2_topk: bad query sizes");
  SB_REQUIRE(db && xn && xn_max && q && out_idx && out_dist && overflow_out, "sb_l2_topk: NULL pointer");
  if (!sb_l2_topk_supported(N, D, ldd, k) || (reinterpret_cast<uintptr_t>(db) & 15u)) {
    sb::set_error("sb_l2_topk: needs D %% 16 == 0, ldd %% 4 == 0, 16-byte aligned db, 1 <= k <= 256, 1 <= N < 2^32");
    return SB_ERR_UNSUPPORTED;
  }
  const L2Plan p = make_l2_plan(D, Q, k);
  if (workspace == nullptr || workspace_bytes < p.total) {
    sb::set_error("sb_l2_topk: workspace too small (%zu < %zu bytes)", workspace_bytes, p.total);
    return SB_ERR_WORKSPACE;
  }
  SB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "sb_l2_topk: workspace must be 256-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  uint32_t* img = reinterpret_cast<uint32_t*>(ws + p.off_img);
  float* qn = reinterpret_cast<float*>(ws + p.off_qn);
  float* tau = reinterpret_cast<float*>(ws + p.off_tau);
  float* margin = reinterpret_cast<float*>(ws + p.off_margin);
  float* tq = reinterpret_cast<float*>(ws + p.off_tq);
  int* cnt = reinterpret_cast<int*>(ws + p.off_cnt);
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(ws + p.off_buf);
  long long* cand_idx = reinterpret_cast<long long*>(ws + p.off_cand_idx);
  long long* cand_off = reinterpret_cast<long long*>(ws + p.off_cand_off);
  long long* cand_cnt = reinterpret_cast<long long*>(ws + p.off_cand_cnt);
  double* dist = reinterpret_cast<double*>(ws + p.off_dist);

  // D <= 128: TMA-fed kernel with the query block resident in shared memory; else the register-staged one
  const bool use_tma = sb::tc_l2_tma_supported(D, ldd, db) != 0;
  {
    if (use_tma) {
      if (int rc = sb::tc_l2_tma_query_image(q, Q, D, ldq, p.col_blocks, img, qn, st)) return rc;
    } else {
      sb::ProfScope prof("l2_query_image_kernel", st);
      query_image_kernel<<<1184, 256, 0, st>>>(q, Q, D, ldq, p.col_blocks, img, qn);
      sb::count_launch();
      if (int rc = sb::check_launch("query_image_kernel")) return rc;
    }
    const float eps = use_tma ? L2_EPS_TMA : (L2_PASSES == 3 ? L2_EPS_3X : L2_EPS_1X);
    l2_init_kernel<<<(p.cols + 255) / 256, 256, 0, st>>>(Q, p.cols, qn, reinterpret_cast<const int*>(xn_max), tau, margin, tq,
                                                        cnt, overflow_out, eps);
    sb::count_launch();
    if (int rc = sb::check_launch("l2_init_kernel")) return rc;
  }
  SB_CUDA_TRY(cudaFuncSetAttribute(l2_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(p.cap * sizeof(unsigned long long))));
  // chunks: the first keeps every row (tau = inf), each later one is GROWTH x the rows seen so far
  int64_t done = 0;
  while (done < N) {
    int64_t len = (done == 0) ? FIRST_CHUNK_MAX : done * (GROWTH - 1);
    if (done == 0 && len > p.cap) len = p.cap;
    if (len > N - done) len = N - done;
    if (done == 0 && use_tma) {
      // seed chunk through the tensor cores in "dense" mode: every pair's key goes to buf[query][row]
      if (int rc = sb::tc_l2_tma_threshold_image(Q, p.cols, D, qn, tq, img, st)) return rc;
      if (int rc = sb::tc_l2_filter_tma(db, len, D, ldd, img, p.col_blocks, xn, tq, buf, cnt, p.cap, 0u, 1, st)) return rc;
      l2_set_count_kernel<<<(Q + 255) / 256, 256, 0, st>>>(cnt, Q, (int)len);
      sb::count_launch();
      if (int rc = sb::check_launch("l2_set_count_kernel")) return rc;
    } else if (done == 0) {
      sb::ProfScope prof("l2_seed_kernel", st);
      l2_seed_kernel<<<Q, 256, D * sizeof(float), st>>>(db, (int)len, D, ldd, q, ldq, buf, cnt, p.cap);
      sb::count_launch();
      if (int rc = sb::check_launch("l2_seed_kernel")) return rc;
    } else if (use_tma) {
      if (int rc = sb::tc_l2_tma_threshold_image(Q, p.cols, D, qn, tq, img, st)) return rc;
      if (int rc = sb::tc_l2_filter_tma(db + done * ldd, len, D, ldd, img, p.col_blocks, xn + done, tq, buf, cnt, p.cap,
                                        (unsigned)done, 0, st))
        return rc;
    } else {
      l2_threshold_image_kernel<<<(p.cols + 255) / 256, 256, 0, st>>>(Q, p.cols, D, qn, tq, img);
      sb::count_launch();
      if (int rc = sb::check_launch("l2_threshold_image_kernel")) return rc;
      if (int rc = sb::tc_l2_filter(db + done * ldd, len, D, ldd, img, p.col_blocks, xn + done, tq, buf, cnt, p.cap,
                                    (unsigned)done, L2_PASSES, st))
        return rc;
    }
    done += len;
    const int final = (done >= N) ? 1 : 0;
    sb::ProfScope prof("l2_compact_kernel", st);
    l2_compact_kernel<<<Q, CP_THREADS, p.cap * sizeof(unsigned long long), st>>>(buf, cnt, p.cap, k, tau, margin, tq,
                                                                                 overflow_out, final, cand_idx, cand_off,
                                                                                 cand_cnt, Q, p.final_pitch);
    sb::count_launch();
    if (int rc = sb::check_launch("l2_compact_kernel")) return rc;
  }
  // exact stage: direct-form distances of the survivors, (distance, row) selection
  if (int rc = sb_rerank_base(db, N, 0, D, ldd, q, Q, ldq, reinterpret_cast<const int64_t*>(cand_idx),
                              reinterpret_cast<const int64_t*>(cand_off), (int64_t)Q * p.final_pitch, SB_METRIC_EUCLIDEAN, dist,
                              stream))
    return rc;
  if (int rc = sb_rerank_select_rows(dist, reinterpret_cast<const int64_t*>(cand_off),
                                     reinterpret_cast<const int64_t*>(cand_cnt), reinterpret_cast<const int64_t*>(cand_idx),
                                     Q, k, 1, out_idx, out_dist, stream))
    return rc;
  return SB_OK;
}

}  // extern "C"
