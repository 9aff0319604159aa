// Stage 3 of the LSH path: candidate re-rank.  Replaces the per-candidate Python
// calls of LSHNearestNeighborIndex._nn (reference: smqtk_indexing/impls/nn_index/
// lsh.py:507-519) into smqtk_indexing/utils/metrics.py:
//   euclidean_distance :83-86      sqrt(sum (a-b)^2)
//   cosine_distance    :120-137    2*acos(clip(a.b/(|a||b|), -1, 1))/pi   (pos_vectors=True)
//   histogram_intersection_distance_fast :70    1 - 0.5*sum(a+b-|a-b|) = 1 - sum(min(a,b))
//
// One warp per candidate row (gather by index), 128-bit loads when rows are
// 16-byte aligned.  Element math is FP32 FFMA with error-free transformations
// (TwoProd via FMA, TwoSum) accumulated as (hi, lo) float pairs -- ~2^-44
// relative, so the ill-conditioned tails (1 - sim near duplicates, 1 - sum(min))
// do not lose the answer; the handful of scalar ops per candidate after the warp
// reduction (sqrt / divide / acos) run in FP64.  Results are written as f64 so
// that ordering ties are not manufactured by rounding to f32.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int RR_WARPS = 8;

struct F2 {  // unevaluated sum hi + lo
  float hi, lo;
};

__device__ __forceinline__ void two_sum_acc(F2& s, float x) {
  // s += x (Knuth TwoSum on hi, error folded into lo)
  float t = s.hi + x;
  float bp = t - s.hi;
  float e = (s.hi - (t - bp)) + (x - bp);
  s.hi = t;
  s.lo += e;
}

__device__ __forceinline__ void acc_prod(F2& s, float a, float b) {
  // s += a*b exactly split into p + e (FMA TwoProd)
  float p = a * b;
  float e = fmaf(a, b, -p);
  two_sum_acc(s, p);
  s.lo += e;
}

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(sb::FULL_MASK, v, o);
  return v;
}

template <int METRIC>
__device__ __forceinline__ void accumulate(F2& s0, F2& s1, F2& s2, float a, float b) {
  if (METRIC == SB_METRIC_EUCLIDEAN) {
    // a - b: exact for near-equal values (Sterbenz), one rounding otherwise; the
    // squared term is split exactly.
    float t = a - b;
    acc_prod(s0, t, t);
  } else if (METRIC == SB_METRIC_COSINE) {
    acc_prod(s0, a, b);
    acc_prod(s1, a, a);
    acc_prod(s2, b, b);
  } else {
    two_sum_acc(s0, fminf(a, b));
  }
}

template <int METRIC>
__global__ void __launch_bounds__(RR_WARPS * 32)
rerank_kernel(const float* __restrict__ db, long long N, int D, long long ldd, const float* __restrict__ q, int Q,
              long long ldq, const long long* __restrict__ cand_idx, const long long* __restrict__ cand_off,
              long long M, double* __restrict__ out, int vec_ok) {
  const int lane = threadIdx.x & 31;
  const long long j = (long long)blockIdx.x * RR_WARPS + (threadIdx.x >> 5);
  if (j >= M) return;
  // owning query: last qi with cand_off[qi] <= j
  int lo = 0, hi = Q;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (cand_off[mid] <= j) lo = mid; else hi = mid;
  }
  const long long row = cand_idx[j];
  if (row < 0 || row >= N) {
    if (lane == 0) out[j] = nan("");
    return;
  }
  const float* a = q + (long long)lo * ldq;
  const float* b = db + row * ldd;
  F2 s0{0.f, 0.f}, s1{0.f, 0.f}, s2{0.f, 0.f};
  if (vec_ok) {
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    for (int i = lane; i < D / 4; i += 32) {
      float4 x = __ldg(a4 + i), y = __ldg(b4 + i);
      accumulate<METRIC>(s0, s1, s2, x.x, y.x);
      accumulate<METRIC>(s0, s1, s2, x.y, y.y);
      accumulate<METRIC>(s0, s1, s2, x.z, y.z);
      accumulate<METRIC>(s0, s1, s2, x.w, y.w);
    }
    for (int i = (D / 4) * 4 + lane; i < D; i += 32) accumulate<METRIC>(s0, s1, s2, __ldg(a + i), __ldg(b + i));
  } else {
    for (int i = lane; i < D; i += 32) accumulate<METRIC>(s0, s1, s2, __ldg(a + i), __ldg(b + i));
  }
  double t0 = warp_sum((double)s0.hi + (double)s0.lo);
  double r;
  if (METRIC == SB_METRIC_EUCLIDEAN) {
    r = sqrt(t0);
  } else if (METRIC == SB_METRIC_COSINE) {
    double t1 = warp_sum((double)s1.hi + (double)s1.lo);
    double t2 = warp_sum((double)s2.hi + (double)s2.lo);
    // metrics.py:111 (scipy cdist 'cosine'): sim = a.b / (|a| |b|); 0/0 -> nan as in the reference
    double sim = t0 / sqrt(t1 * t2);  // sqrt(t*t) == t: identical vectors give exactly 1
    // numpy's minimum/maximum propagate NaN (metrics.py:135); CUDA's fmin/fmax do not
    r = (sim != sim) ? sim : 2.0 * acos(fmax(fmin(sim, 1.0), -1.0)) / 3.14159265358979323846;
  } else {
    r = 1.0 - t0;
  }
  if (lane == 0) out[j] = r;
}

// ---- per-query ordering of the candidates: (distance, position) ascending ----
constexpr int SEL_THREADS = 256;
constexpr int SEL_CAP = 4096;

__device__ __forceinline__ bool before(double da, long long ia, double db_, long long ib) {
  // NaN sorts last (the reference's sorted() leaves NaN order unspecified)
  const bool na = isnan(da), nb = isnan(db_);
  if (na != nb) return nb;
  if (!na && da != db_) return da < db_;
  return ia < ib;
}

__global__ void __launch_bounds__(SEL_THREADS)
rerank_select_kernel(const double* __restrict__ dist, const long long* __restrict__ cand_off, int Q, int n,
                     long long* __restrict__ out_pos, double* __restrict__ out_dist) {
  __shared__ double s_d[SEL_CAP];
  const int qi = blockIdx.x, tid = threadIdx.x;
  const long long beg = cand_off[qi], end = cand_off[qi + 1];
  const long long m = end - beg;
  for (int i = tid; i < n; i += SEL_THREADS) {
    out_pos[(long long)qi * n + i] = -1;
    out_dist[(long long)qi * n + i] = nan("");
  }
  const bool staged = (m <= SEL_CAP);
  if (staged)
    for (long long i = tid; i < m; i += SEL_THREADS) s_d[i] = dist[beg + i];
  __syncthreads();
  for (long long i = tid; i < m; i += SEL_THREADS) {
    const double di = staged ? s_d[i] : dist[beg + i];
    long long rank = 0;
    if (staged) {
      for (long long jx = 0; jx < m; ++jx) rank += before(s_d[jx], jx, di, i) ? 1 : 0;
    } else {
      for (long long jx = 0; jx < m; ++jx) rank += before(dist[beg + jx], jx, di, i) ? 1 : 0;
    }
    if (rank < n) {
      out_pos[(long long)qi * n + rank] = beg + i;
      out_dist[(long long)qi * n + rank] = di;
    }
  }
}

}  // namespace

extern "C" {

int sb_rerank(const float* db, int64_t N, int32_t D, int64_t ldd, const float* q, int32_t Q, int64_t ldq,
              const int64_t* cand_idx, const int64_t* cand_off, int64_t M, int32_t metric, double* out,
              void* stream) {
  SB_REQUIRE(D >= 1 && Q >= 1 && M >= 0 && N >= 0, "sb_rerank: bad sizes");
  SB_REQUIRE(ldd >= D && ldq >= D, "sb_rerank: leading dimension smaller than D");
  SB_REQUIRE(metric >= SB_METRIC_EUCLIDEAN && metric <= SB_METRIC_HIK, "sb_rerank: bad metric %d", metric);
  if (M == 0) return SB_OK;
  SB_REQUIRE(db && q && cand_idx && cand_off && out, "sb_rerank: NULL pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int vec_ok = ((reinterpret_cast<uintptr_t>(db) | reinterpret_cast<uintptr_t>(q)) % 16 == 0 && ldd % 4 == 0 &&
                      ldq % 4 == 0) ? 1 : 0;
  const unsigned grid = (unsigned)((M + RR_WARPS - 1) / RR_WARPS);
  const long long* ci = reinterpret_cast<const long long*>(cand_idx);
  const long long* co = reinterpret_cast<const long long*>(cand_off);
  sb::ProfScope prof("rerank_kernel", st);
  switch (metric) {
    case SB_METRIC_EUCLIDEAN:
      rerank_kernel<SB_METRIC_EUCLIDEAN><<<grid, RR_WARPS * 32, 0, st>>>(db, N, D, ldd, q, Q, ldq, ci, co, M, out, vec_ok);
      break;
    case SB_METRIC_COSINE:
      rerank_kernel<SB_METRIC_COSINE><<<grid, RR_WARPS * 32, 0, st>>>(db, N, D, ldd, q, Q, ldq, ci, co, M, out, vec_ok);
      break;
    default:
      rerank_kernel<SB_METRIC_HIK><<<grid, RR_WARPS * 32, 0, st>>>(db, N, D, ldd, q, Q, ldq, ci, co, M, out, vec_ok);
      break;
  }
  sb::count_launch();
  return sb::check_launch("rerank_kernel");
}

int sb_rerank_select(const double* dist, const int64_t* cand_off, int32_t Q, int32_t n, int64_t* out_pos,
                     double* out_dist, void* stream) {
  SB_REQUIRE(Q >= 1 && n >= 1, "sb_rerank_select: bad sizes");
  SB_REQUIRE(cand_off && out_pos && out_dist, "sb_rerank_select: NULL pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  sb::ProfScope prof("rerank_select_kernel", st);
  rerank_select_kernel<<<Q, SEL_THREADS, 0, st>>>(dist, reinterpret_cast<const long long*>(cand_off), Q, n,
                                                  reinterpret_cast<long long*>(out_pos), out_dist);
  sb::count_launch();
  return sb::check_launch("rerank_select_kernel");
}

}  // extern "C"
