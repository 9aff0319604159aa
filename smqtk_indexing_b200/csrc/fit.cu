// Training contractions of ItqFunctor.fit on the GPU (reference:
// smqtk_indexing/impls/lsh_functor/itq.py:338-383 and _find_itq_rotation :239-289):
//   row divisors (normalisation)           itq.py:340, 184-189
//   column means                           itq.py:343
//   covariance  Xc^T Xc                    itq.py:351   (np.cov, float64)
//   projection  V = Xc . pc_top            itq.py:378
//   per ITQ iteration  Z = V.R, sign, C = UX^T V     itq.py:270-274
// The small dense factorizations (eig of DxD, SVD of bxb) stay on host LAPACK
// like the reference.  Everything here accumulates in FP64 (the reference's
// dtype), FP64 DFMA on CUDA cores; reductions over rows are two-stage with a
// fixed summation order, so results are run-to-run deterministic.
//
// Operands are described by SbOperand so one GEMM pair covers all cases:
//   kind 0: f32 matrix, value = x / div[r] - mean[c]   (div / mean optional)
//   kind 1: f64 matrix, value = x / div[r] - mean[c]
//   kind 2: packed sign bits (uint32[n][W], b bits): value = bit ? +1 : -1
#include "common.cuh"

namespace {

struct Operand {
  const void* p;
  long long ld;        // row pitch in elements (words for kind 2)
  int cols;
  int kind;
  const double* div;   // [n] or null
  const double* mean;  // [cols] or null
  int bits;            // kind 2: code length b
};

__device__ __forceinline__ double fetch(const Operand& o, long long r, int c) {
  if (o.kind == 2) {
    const uint32_t* row = static_cast<const uint32_t*>(o.p) + r * o.ld;
    const int pbit = o.bits - 1 - c;
    const int w = (int)o.ld - 1 - pbit / 32;
    return ((row[w] >> (pbit & 31)) & 1u) ? 1.0 : -1.0;
  }
  double v = (o.kind == 0) ? (double)static_cast<const float*>(o.p)[r * o.ld + c]
                           : static_cast<const double*>(o.p)[r * o.ld + c];
  if (o.div) v = v / o.div[r];
  if (o.mean) v -= o.mean[c];
  return v;
}

// ---- row divisors -----------------------------------------------------------------
__global__ void row_div_kernel(Operand x, long long n, int norm_kind, double p, double* __restrict__ div) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  double acc = 0.0;
  for (int c = lane; c < x.cols; c += 32) {
    double a = fabs(fetch(x, r, c));
    if (norm_kind == SB_NORM_LP) acc += (p == 2.0) ? a * a : (p == 1.0 ? a : pow(a, p));
    else if (norm_kind == SB_NORM_INF) acc = fmax(acc, a);
    else acc += (a != 0.0) ? 1.0 : 0.0;
  }
  for (int o = 16; o; o >>= 1) {
    double t = __shfl_xor_sync(sb::FULL_MASK, acc, o);
    acc = (norm_kind == SB_NORM_INF) ? fmax(acc, t) : acc + t;
  }
  if (lane == 0) {
    double nrm = acc;
    if (norm_kind == SB_NORM_LP) nrm = (p == 2.0) ? sqrt(acc) : (p == 1.0 ? acc : pow(acc, 1.0 / p));
    div[r] = (nrm == 0.0) ? 1.0 : nrm;
  }
}

// ---- column sums (two stage, fixed order) ------------------------------------------
__global__ void col_sum_partial_kernel(Operand x, long long n, long long rows_per_block, double* __restrict__ partial) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= x.cols) return;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(n, r0 + rows_per_block);
  double acc = 0.0;
  for (long long r = r0; r < r1; ++r) acc += fetch(x, r, c);
  partial[(size_t)blockIdx.y * x.cols + c] = acc;
}

__global__ void reduce_partials_kernel(const double* __restrict__ partial, int parts, long long elems, double scale,
                                       double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= elems) return;
  double acc = 0.0;
  for (int p = 0; p < parts; ++p) acc += partial[(size_t)p * elems + i];
  out[i] = acc * scale;
}

// ---- FP64 register-tiled micro-kernel shared by the two GEMMs --------------------------------
// CTA tile 128 x 128, 256 threads, 8 x 8 outputs per thread, K slab of 8 staged in shared memory.
// A thread owns output rows {2*ty, 2*ty+1} + 32*m and columns {2*tx, 2*tx+1} + 32*m (m = 0..3):
// every shared-memory read is a 128-bit LDS whose 16 lanes of a half-warp touch 256 contiguous
// bytes (conflict-free), and one K step costs 8 LDS.128 for 64 DFMA -- the DFMA pipe (64 lanes/clk/SM)
// and the shared-memory pipe finish together.
constexpr int GT = 128, GK = 8;

__device__ __forceinline__ void dgemm_slab(const double (&As)[GK][GT], const double (&Bs)[GK][GT], int ty, int tx,
                                           double (&acc)[8][8]) {
#pragma unroll
  for (int kk = 0; kk < GK; ++kk) {
    double a[8], b[8];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const double2 av = *reinterpret_cast<const double2*>(&As[kk][32 * m + 2 * ty]);
      const double2 bv = *reinterpret_cast<const double2*>(&Bs[kk][32 * m + 2 * tx]);
      a[2 * m] = av.x; a[2 * m + 1] = av.y;
      b[2 * m] = bv.x; b[2 * m + 1] = bv.y;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
  }
}
// output index of accumulator slot i (rows) / j (columns) of thread coordinate t
__device__ __forceinline__ int tile_index(int t, int i) { return 32 * (i >> 1) + 2 * t + (i & 1); }

// ---- gram: C[Ma][Mb] = sum_r A[r][:]^T B[r][:] ---------------------------------------
__global__ void __launch_bounds__(256)
gram_partial_kernel(Operand A, Operand B, long long n, long long rows_per_block, double* __restrict__ partial) {
  __shared__ __align__(16) double As[GK][GT];
  __shared__ __align__(16) double Bs[GK][GT];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int a0 = blockIdx.x * GT, b0 = blockIdx.y * GT;
  const long long r0 = (long long)blockIdx.z * rows_per_block;
  const long long r1 = min(n, r0 + rows_per_block);
  double acc[8][8] = {};
  // staging: thread -> (slab row tid / 32, 4 columns (tid % 32) * 4 ..)
  const int lk = tid >> 5;
  const int lc = (tid & 31) * 4;
  // register prefetch: the next slab's global loads are in flight while this slab is multiplied
  double pa[4], pb[4];
  auto load_slab = [&](long long rr) {
    const long long r = rr + lk;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ca = a0 + lc + j, cb = b0 + lc + j;
      pa[j] = (r < r1 && ca < A.cols) ? fetch(A, r, ca) : 0.0;
      pb[j] = (r < r1 && cb < B.cols) ? fetch(B, r, cb) : 0.0;
    }
  };
  if (r0 < r1) load_slab(r0);
  for (long long rr = r0; rr < r1; rr += GK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[lk][lc + j] = pa[j];
      Bs[lk][lc + j] = pb[j];
    }
    __syncthreads();
    if (rr + GK < r1) load_slab(rr + GK);
    dgemm_slab(As, Bs, ty, tx, acc);
    __syncthreads();
  }
  double* out = partial + (size_t)blockIdx.z * A.cols * B.cols;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ra = a0 + tile_index(ty, i), cb = b0 + tile_index(tx, j);
      if (ra < A.cols && cb < B.cols) out[(size_t)ra * B.cols + cb] = acc[i][j];
    }
}

// ---- tall GEMM: out[n][M] = A[n][K] . Bm[K][M], optional sign/pack epilogue -----------
__global__ void __launch_bounds__(256)
project_kernel(Operand A, const double* __restrict__ Bm, int M, long long n, double* __restrict__ out_f64,
               uint32_t* __restrict__ out_codes, int Wc) {
  __shared__ __align__(16) double As[GK][GT];       // As[k][row]
  __shared__ __align__(16) double Bs[GK][GT];       // Bs[k][col]
  __shared__ uint32_t s_bits[GT][5];                // a 128-column tile touches at most 5 words
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const long long row0 = (long long)blockIdx.x * GT;
  const int col0 = blockIdx.y * GT;
  const int K = A.cols;
  double acc[8][8] = {};
  const int ar = tid >> 1, ak = (tid & 1) * 4;    // A tile: 128 rows x 8 k, 4 consecutive k per thread
  const int bk = tid >> 5, bc = (tid & 31) * 4;   // B tile: 8 k x 128 cols
  for (int i = tid; i < GT * 5; i += 256) (&s_bits[0][0])[i] = 0u;
  double pa[4], pb[4];
  auto load_slab = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = k0 + ak + i;
      pa[i] = (row0 + ar < n && kk < K) ? fetch(A, row0 + ar, kk) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = k0 + bk, cc = col0 + bc + j;
      pb[j] = (kk < K && cc < M) ? Bm[(size_t)kk * M + cc] : 0.0;
    }
  };
  load_slab(0);
  for (int k0 = 0; k0 < K; k0 += GK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) As[ak + i][ar] = pa[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) Bs[bk][bc + j] = pb[j];
    __syncthreads();
    if (k0 + GK < K) load_slab(k0 + GK);
    dgemm_slab(As, Bs, ty, tx, acc);
    __syncthreads();
  }
  const int w_first = out_codes ? (Wc - 1 - (M - 1 - col0) / 32) : 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = tile_index(ty, i);
    const long long row = row0 + r;
    if (row >= n) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = col0 + tile_index(tx, j);
      if (col >= M) continue;
      const double z = acc[i][j];
      if (out_f64) out_f64[row * (long long)M + col] = z;
      if (out_codes && z >= 0.0) {
        const int pbit = M - 1 - col;
        atomicOr(&s_bits[r][(Wc - 1 - pbit / 32) - w_first], 1u << (pbit & 31));
      }
    }
  }
  if (out_codes) {
    __syncthreads();
    for (int i = tid; i < GT * 5; i += 256) {
      const int r = i / 5, wi = i - r * 5;
      const uint32_t v = s_bits[r][wi];
      if (row0 + r < n && v != 0u && w_first + wi < Wc) atomicOr(&out_codes[(row0 + r) * Wc + w_first + wi], v);
    }
  }
}

int make_operand(Operand& o, const void* p, int kind, long long ld, int cols, const double* div, const double* mean,
                 int bits) {
  SB_REQUIRE(kind >= 0 && kind <= 2, "fit operand: bad kind %d", kind);
  SB_REQUIRE(p != nullptr && cols >= 1, "fit operand: NULL pointer or no columns");
  o.p = p; o.ld = ld; o.cols = cols; o.kind = kind; o.div = div; o.mean = mean; o.bits = bits;
  return SB_OK;
}

int gram_chunks(long long n, int Ma, int Mb) {
  const long long tiles = (long long)((Ma + GT - 1) / GT) * ((Mb + GT - 1) / GT);
  long long chunks = (2ll * sb::sm_count() + tiles - 1) / tiles;
  const long long max_chunks = (n + 255) / 256;
  return (int)max(1ll, min(chunks, max(1ll, max_chunks)));
}

}  // namespace

extern "C" {

int sb_fit_row_div(const void* X, int32_t x_kind, int64_t n, int32_t D, int64_t ldx, int32_t norm_kind,
                          double norm_p, double* div_out, void* stream) {
  SB_REQUIRE(norm_kind >= SB_NORM_LP && norm_kind <= SB_NORM_L0, "sb_fit_row_div: bad norm_kind %d", norm_kind);
  SB_REQUIRE(x_kind == 0 || x_kind == 1, "sb_fit_row_div: X must be f32 (0) or f64 (1)");
  if (n == 0) return SB_OK;
  Operand x;
  if (int rc = make_operand(x, X, x_kind, ldx, D, nullptr, nullptr, 0)) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  sb::ProfScope prof("row_div_kernel", st);
  row_div_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(x, n, norm_kind, norm_p, div_out);
  sb::count_launch();
  return sb::check_launch("row_div_kernel");
}

// mean_out[c] = (1/n) sum_r (X[r][c] / div[r]);  workspace: sb_fit_workspace_bytes
size_t sb_fit_workspace_bytes(int64_t n, int32_t Ma, int32_t Mb) {
  const int chunks = gram_chunks(n, Ma, Mb);
  const size_t gram = (size_t)chunks * Ma * Mb * sizeof(double);
  const size_t cols = (size_t)256 * (Ma > Mb ? Ma : Mb) * sizeof(double);
  return gram > cols ? gram : cols;
}

int sb_fit_col_mean(const void* X, int32_t x_kind, int64_t n, int32_t D, int64_t ldx, const double* div,
                           double* mean_out, void* workspace, size_t workspace_bytes, void* stream) {
  SB_REQUIRE(n >= 1 && D >= 1, "sb_fit_col_mean: empty input");
  SB_REQUIRE(x_kind == 0 || x_kind == 1, "sb_fit_col_mean: X must be f32 (0) or f64 (1)");
  Operand x;
  if (int rc = make_operand(x, X, x_kind, ldx, D, div, nullptr, 0)) return rc;
  const int parts = (int)min(256ll, (long long)((n + 255) / 256));
  if (workspace_bytes < (size_t)parts * D * sizeof(double)) {
    sb::set_error("sb_fit_col_mean: workspace too small");
    return SB_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(workspace);
  const long long rpb = (n + parts - 1) / parts;
  dim3 grid((D + 127) / 128, parts);
  sb::ProfScope prof("col_mean_kernels", st);
  col_sum_partial_kernel<<<grid, 128, 0, st>>>(x, n, rpb, partial);
  reduce_partials_kernel<<<(D + 255) / 256, 256, 0, st>>>(partial, parts, D, 1.0 / (double)n, mean_out);
  sb::count_launch(2);
  return sb::check_launch("sb_fit_col_mean");
}

// out[Ma][Mb] = scale * sum_r opA(r)^T opB(r)
int sb_fit_gram(const void* A, int32_t a_kind, int64_t lda, int32_t Ma, const double* a_div,
                       const double* a_mean, int32_t a_bits, const void* B, int32_t b_kind, int64_t ldb, int32_t Mb,
                       const double* b_div, const double* b_mean, int32_t b_bits, int64_t n, double scale,
                       double* out, void* workspace, size_t workspace_bytes, void* stream) {
  SB_REQUIRE(n >= 1, "sb_fit_gram: no rows");
  Operand a, b;
  if (int rc = make_operand(a, A, a_kind, lda, Ma, a_div, a_mean, a_bits)) return rc;
  if (int rc = make_operand(b, B, b_kind, ldb, Mb, b_div, b_mean, b_bits)) return rc;
  const int chunks = gram_chunks(n, Ma, Mb);
  if (workspace_bytes < (size_t)chunks * Ma * Mb * sizeof(double)) {
    sb::set_error("sb_fit_gram: workspace too small");
    return SB_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(workspace);
  long long rpb = (n + chunks - 1) / chunks;
  rpb = ((rpb + GK - 1) / GK) * GK;
  dim3 grid((Ma + GT - 1) / GT, (Mb + GT - 1) / GT, chunks);
  sb::ProfScope prof("gram_kernels", st);
  gram_partial_kernel<<<grid, 256, 0, st>>>(a, b, n, rpb, partial);
  const long long elems = (long long)Ma * Mb;
  reduce_partials_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(partial, chunks, elems, scale, out);
  sb::count_launch(2);
  return sb::check_launch("sb_fit_gram");
}

// out = opA[n][K] . Bm[K][M]; out_f64 (f64[n][M]) and/or out_codes (u32[n][Wc], sign bits) may be NULL
int sb_fit_project(const void* A, int32_t a_kind, int64_t lda, int32_t K, const double* a_div,
                          const double* a_mean, const double* Bm, int32_t M, int64_t n, double* out_f64,
                          uint32_t* out_codes, int32_t Wc, void* stream) {
  SB_REQUIRE(a_kind == 0 || a_kind == 1, "sb_fit_project: A must be f32 (0) or f64 (1)");
  SB_REQUIRE(Bm != nullptr && M >= 1, "sb_fit_project: bad B");
  SB_REQUIRE(out_codes == nullptr || Wc * 32 >= M, "sb_fit_project: %d words cannot hold %d bits", Wc, M);
  if (n == 0) return SB_OK;
  Operand a;
  if (int rc = make_operand(a, A, a_kind, lda, K, a_div, a_mean, 0)) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (out_codes) SB_CUDA_TRY(cudaMemsetAsync(out_codes, 0, (size_t)n * Wc * sizeof(uint32_t), st));
  dim3 grid((unsigned)((n + GT - 1) / GT), (M + GT - 1) / GT);
  sb::ProfScope prof("project_kernel", st);
  project_kernel<<<grid, 256, 0, st>>>(a, Bm, M, n, out_f64, out_codes, Wc);
  sb::count_launch();
  return sb::check_launch("project_kernel");
}

}  // extern "C"
