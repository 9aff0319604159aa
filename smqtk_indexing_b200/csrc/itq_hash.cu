// Stage 1 of the LSH path, shape-generic kernel: ITQ hashing
//   codes[r] = pack( ((X[r] / norm(X[r])) - mean) . R  >= 0 )
// Replaces ItqFunctor.get_hash (reference: smqtk_indexing/impls/lsh_functor/
// itq.py:389-408, normalisation :172-191) and the bit-vector -> int step that
// follows it everywhere (smqtk_indexing/utils/bits.py:4-20).
//
// This file holds the FP32 FFMA kernel that accepts ANY (n, D, b): a register-
// tiled SGEMM (64x64 CTA tile, 4x4 per thread, K staged through shared memory)
// with everything else fused: row norms and centring are applied while the A
// tile is staged, the sign test and the big-endian bit-pack happen in the
// epilogue (shared-memory OR, then one global atomicOr per row/word/CTA).
// The tcgen05 3xTF32 tensor-core kernel for aligned shapes lives in
// itq_hash_tc.cu (sb_itq_hash_tc); it needs the rotation pre-split into its
// shared-memory image, so it has its own entry points.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int THREADS = 256;

// 1 / norm(row) folded into a divisor: returns the value the row is divided by
// (1 when normalisation is off or the norm is 0: itq.py:186-188).
__device__ __forceinline__ float finish_norm(float acc, int kind, float p) {
  float n;
  if (kind == SB_NORM_LP) {
    if (p == 2.0f) n = sqrtf(acc);
    else if (p == 1.0f) n = acc;
    else n = powf(acc, 1.0f / p);
  } else {
    n = acc;  // INF: max |x| ; L0: count
  }
  return (n == 0.0f) ? 1.0f : n;
}

__device__ __forceinline__ float norm_term(float x, int kind, float p) {
  float a = fabsf(x);
  if (kind == SB_NORM_LP) {
    if (p == 2.0f) return a * a;
    if (p == 1.0f) return a;
    return powf(a, p);
  }
  if (kind == SB_NORM_L0) return (x != 0.0f) ? 1.0f : 0.0f;
  return a;  // INF handled with max
}

__global__ void __launch_bounds__(THREADS)
itq_hash_simt_kernel(const float* __restrict__ X, long long n, int D, long long ldx, const float* __restrict__ mean,
                     const float* __restrict__ R, int b, int norm_kind, float norm_p, uint32_t* __restrict__ codes,
                     int W, float* __restrict__ z_out) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN];
  __shared__ float s_div[BM];
  __shared__ uint32_t s_bits[BM][3];  // a 64-column tile touches at most 3 words

  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;

  // ---- row divisors (normalisation), 4 threads per row ----
  {
    const int r = tid >> 2, part = tid & 3;
    const long long row = row0 + r;
    float acc = 0.0f;
    if (norm_kind != SB_NORM_NONE && row < n) {
      const float* xr = X + row * ldx;
      for (int d = part; d < D; d += 4) {
        float t = norm_term(xr[d], norm_kind, norm_p);
        acc = (norm_kind == SB_NORM_INF) ? fmaxf(acc, t) : acc + t;
      }
    }
    float o1 = __shfl_xor_sync(sb::FULL_MASK, acc, 1);
    acc = (norm_kind == SB_NORM_INF) ? fmaxf(acc, o1) : acc + o1;
    float o2 = __shfl_xor_sync(sb::FULL_MASK, acc, 2);
    acc = (norm_kind == SB_NORM_INF) ? fmaxf(acc, o2) : acc + o2;
    if (part == 0) s_div[r] = (norm_kind == SB_NORM_NONE) ? 1.0f : finish_norm(acc, norm_kind, norm_p);
    if (tid < BM * 3) (&s_bits[0][0])[tid] = 0u;
  }
  __syncthreads();

  const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  // staging maps: A tile 64 rows x 16 k -> thread (ar = tid/4, ak = (tid%4)*4 .. +3)
  //               B tile 16 k x 64 cols -> thread (bk = tid/16, bc = (tid%16)*4 .. +3)
  const int ar = tid >> 2, ak = (tid & 3) * 4;
  const int bk = tid >> 4, bc = (tid & 15) * 4;
  const long long arow = row0 + ar;
  const float adiv = s_div[ar];

  for (int k0 = 0; k0 < D; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int kk = k0 + ak + i;
      float v = 0.0f;
      if (arow < n && kk < D) {
        float x = X[arow * ldx + kk];
        if (norm_kind != SB_NORM_NONE) x = x / adiv;
        v = x - (mean ? mean[kk] : 0.0f);
      }
      As[ak + i][ar] = v;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kk = k0 + bk, cc = col0 + bc + j;
      Bs[bk][bc + j] = (kk < D && cc < b) ? R[(size_t)kk * b + cc] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: sign -> bit (column j is integer bit p = b-1-j) ----
  const int w_first = W - 1 - (b - 1 - col0) / 32;  // word holding the tile's first column
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    const long long row = row0 + r;
    if (row >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      if (col >= b) continue;
      const float z = acc[i][j];
      if (z_out) z_out[row * (long long)b + col] = z;
      if (z >= 0.0f) {
        const int p = b - 1 - col;
        const int w = W - 1 - p / 32;
        atomicOr(&s_bits[r][w - w_first], 1u << (p & 31));
      }
    }
  }
  __syncthreads();
  if (tid < BM * 3) {
    const int r = tid / 3, wi = tid - r * 3;
    const long long row = row0 + r;
    const uint32_t v = s_bits[r][wi];
    const int w = w_first + wi;
    if (row < n && v != 0u && w < W) atomicOr(&codes[row * W + w], v);
  }
}

// One warp per row: div[r] = norm(X[r]) (1 when 0), the divisor of itq.py:184-189.
__global__ void __launch_bounds__(256)
row_div_f32_kernel(const float* __restrict__ X, long long n, int D, long long ldx, int norm_kind, float norm_p,
                   float* __restrict__ div_out) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xr = X + row * ldx;
  float acc = 0.0f;
  for (int d = lane; d < D; d += 32) {
    const float t = norm_term(xr[d], norm_kind, norm_p);
    acc = (norm_kind == SB_NORM_INF) ? fmaxf(acc, t) : acc + t;
  }
  for (int o = 16; o; o >>= 1) {
    const float t = __shfl_xor_sync(sb::FULL_MASK, acc, o);
    acc = (norm_kind == SB_NORM_INF) ? fmaxf(acc, t) : acc + t;
  }
  if (lane == 0) div_out[row] = finish_norm(acc, norm_kind, norm_p);
}

}  // namespace

extern "C" int sb_itq_row_div(const float* X, int64_t n, int32_t D, int64_t ldx, int32_t norm_kind, float norm_p,
                              float* div_out, void* stream) {
  SB_REQUIRE(n >= 0 && D >= 1 && ldx >= D, "sb_itq_row_div: bad sizes");
  SB_REQUIRE(norm_kind >= SB_NORM_LP && norm_kind <= SB_NORM_L0, "sb_itq_row_div: bad norm_kind %d", norm_kind);
  SB_REQUIRE(norm_kind != SB_NORM_LP || norm_p > 0.0f, "sb_itq_row_div: Lp norm needs p > 0");
  if (n == 0) return SB_OK;
  SB_REQUIRE(X != nullptr && div_out != nullptr, "sb_itq_row_div: NULL pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  sb::ProfScope prof("row_div_f32_kernel", st);
  row_div_f32_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(X, n, D, ldx, norm_kind, norm_p, div_out);
  sb::count_launch();
  return sb::check_launch("row_div_f32_kernel");
}

extern "C" int sb_itq_hash(const float* X, int64_t n, int32_t D, int64_t ldx, const float* mean, const float* R,
                           int32_t b, int32_t norm_kind, float norm_p, uint32_t* codes_out, int32_t W, float* z_out,
                           int32_t variant, void* stream) {
  SB_REQUIRE(n >= 0 && D >= 1 && b >= 1, "sb_itq_hash: need n>=0, D>=1, b>=1 (n=%lld D=%d b=%d)", (long long)n, D, b);
  SB_REQUIRE(ldx >= D, "sb_itq_hash: ldx=%lld < D=%d", (long long)ldx, D);
  SB_REQUIRE(W * 32 >= b, "sb_itq_hash: %d words cannot hold %d bits", W, b);
  SB_REQUIRE(X != nullptr || n == 0, "sb_itq_hash: X is NULL");
  SB_REQUIRE(R != nullptr && codes_out != nullptr, "sb_itq_hash: NULL pointer");
  SB_REQUIRE(norm_kind >= SB_NORM_NONE && norm_kind <= SB_NORM_L0, "sb_itq_hash: bad norm_kind %d", norm_kind);
  SB_REQUIRE(norm_kind != SB_NORM_LP || norm_p > 0.0f, "sb_itq_hash: Lp norm needs p > 0");
  SB_REQUIRE(variant == 0 || variant == 1, "sb_itq_hash: bad variant %d (the tensor-core kernel is sb_itq_hash_tc)", variant);
  if (n == 0) return SB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  SB_CUDA_TRY(cudaMemsetAsync(codes_out, 0, (size_t)n * W * sizeof(uint32_t), st));
  dim3 grid((unsigned)((n + BM - 1) / BM), (unsigned)((b + BN - 1) / BN));
  sb::ProfScope prof("itq_hash_simt_kernel", st);
  itq_hash_simt_kernel<<<grid, THREADS, 0, st>>>(X, n, D, ldx, mean, R, b, norm_kind, norm_p, codes_out, W, z_out);
  sb::count_launch();
  return sb::check_launch("itq_hash_simt_kernel");
}
