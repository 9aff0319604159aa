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
    // a - b = t + e exactly (TwoSum), (t + e)^2 = t^2 + 2 t e + O(e^2): t^2 is split
    // exactly, the cross term goes to the low word -- ~2^-45 relative per term.
    const float t = a - b;
    const float bv = t - a;
    const float e = (a - (t - bv)) + (-b - bv);
    acc_prod(s0, t, t);
    s0.lo = fmaf(2.0f * t, e, s0.lo);
  } else if (METRIC == SB_METRIC_COSINE) {
    acc_prod(s0, a, b);
    acc_prod(s1, a, a);
    acc_prod(s2, b, b);
  } else {
    two_sum_acc(s0, fminf(a, b));
  }
}

template <int METRIC>
__device__ __forceinline__ void accumulate4(F2& s0, F2& s1, F2& s2, const float4& x, const float4& y) {
  accumulate<METRIC>(s0, s1, s2, x.x, y.x);
  accumulate<METRIC>(s0, s1, s2, x.y, y.y);
  accumulate<METRIC>(s0, s1, s2, x.z, y.z);
  accumulate<METRIC>(s0, s1, s2, x.w, y.w);
}

// candidate rows are read exactly once: keep them out of L1 (the query row, shared by the query's candidates, stays)
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

template <int METRIC>
__global__ void __launch_bounds__(RR_WARPS * 32)
rerank_kernel(const float* __restrict__ db, long long N, int D, long long ldd, const float* __restrict__ q, int Q,
              long long ldq, const long long* __restrict__ cand_idx, const long long* __restrict__ cand_off,
              long long M, double* __restrict__ out, int vec_ok, long long row_base, double miss_value,
              const float* const* __restrict__ shards, const long long* __restrict__ shard_bounds, int n_shards,
              long long pitch) {
  const int lane = threadIdx.x & 31;
  const long long j = (long long)blockIdx.x * RR_WARPS + (threadIdx.x >> 5);
  if (j >= M) return;
  // owning query: fixed-pitch segments (sb_expand_candidates) need one division; ragged lists a binary
  // search for the last qi with cand_off[qi] <= j (12 dependent L2 loads at Q = 4096: as long as the row itself)
  int lo = 0;
  if (pitch > 0) {
    lo = (int)(j / pitch);
  } else {
    int hi = Q;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (cand_off[mid] <= j) lo = mid; else hi = mid;
    }
  }
  long long row = cand_idx[j] - row_base;                    // db holds global rows [row_base, row_base + N)
  const float* b;
  if (shards != nullptr) {
    // row-sharded table: shard s holds global rows [bounds[s], bounds[s+1]); the other shards are peer
    // GPUs' HBM mapped into this process (CUDA IPC) and read over NVLink -- no collective in the re-rank
    int sh = -1;
    if (row >= 0)
      for (int t = 0; t < n_shards; ++t)
        if (row >= shard_bounds[t] && row < shard_bounds[t + 1]) sh = t;
    if (sh < 0) {
      if (lane == 0) out[j] = miss_value;
      return;
    }
    b = shards[sh] + (row - shard_bounds[sh]) * ldd;
  } else {
    if (row < 0 || row >= N) {
      if (lane == 0) out[j] = miss_value;
      return;
    }
    b = db + row * ldd;
  }
  const float* a = q + (long long)lo * ldq;
  F2 s0{0.f, 0.f}, s1{0.f, 0.f}, s2{0.f, 0.f};
  if (vec_ok) {
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    const int n4 = D / 4;
    int i = lane;
    // a gathered row is read once, by one warp: what bounds this kernel is memory LATENCY, so four 128-bit
    // loads per operand are issued before any of them is consumed (a 2 KB row = 4 loads per lane in flight)
    for (; i + 96 < n4; i += 128) {
      float4 y0 = ldg_stream4(b4 + i), y1 = ldg_stream4(b4 + i + 32), y2 = ldg_stream4(b4 + i + 64), y3 = ldg_stream4(b4 + i + 96);
      float4 x0 = __ldg(a4 + i), x1 = __ldg(a4 + i + 32), x2 = __ldg(a4 + i + 64), x3 = __ldg(a4 + i + 96);
      accumulate4<METRIC>(s0, s1, s2, x0, y0);
      accumulate4<METRIC>(s0, s1, s2, x1, y1);
      accumulate4<METRIC>(s0, s1, s2, x2, y2);
      accumulate4<METRIC>(s0, s1, s2, x3, y3);
    }
    for (; i < n4; i += 32) {
      float4 x = __ldg(a4 + i), y = ldg_stream4(b4 + i);
      accumulate4<METRIC>(s0, s1, s2, x, y);
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

// cand_cnt (optional): candidates actually present in the query's segment (fixed-pitch
// layout of sb_expand_candidates); cand_idx (optional): write candidate ROWS instead of
// positions.
// tie_by_row: equal distances are ordered by candidate ROW (cand_idx value) instead of position.
__global__ void __launch_bounds__(SEL_THREADS)
rerank_select_kernel(const double* __restrict__ dist, const long long* __restrict__ cand_off,
                     const long long* __restrict__ cand_cnt, const long long* __restrict__ cand_idx, int Q, int n,
                     long long* __restrict__ out_pos, double* __restrict__ out_dist, int tie_by_row) {
  __shared__ double s_d[SEL_CAP];
  const int qi = blockIdx.x, tid = threadIdx.x;
  const long long beg = cand_off[qi], end = cand_off[qi + 1];
  const long long m = cand_cnt ? min(cand_cnt[qi], end - beg) : end - beg;
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
    if (tie_by_row) {
      const long long ri = cand_idx[beg + i];
      for (long long jx = 0; jx < m; ++jx)
        rank += before(staged ? s_d[jx] : dist[beg + jx], cand_idx[beg + jx], di, ri) ? 1 : 0;
    } else if (staged) {
      for (long long jx = 0; jx < m; ++jx) rank += before(s_d[jx], jx, di, i) ? 1 : 0;
    } else {
      for (long long jx = 0; jx < m; ++jx) rank += before(dist[beg + jx], jx, di, i) ? 1 : 0;
    }
    if (rank < n) {
      out_pos[(long long)qi * n + rank] = cand_idx ? cand_idx[beg + i] : beg + i;
      out_dist[(long long)qi * n + rank] = di;
    }
  }
}

// ---- candidate expansion: near codes -> descriptor rows (lsh.py:490-496) ------------
// One CTA per query.  Segment q of cand_idx is [q*pitch, (q+1)*pitch): the rows of the
// query's near codes in (code rank, row) order, then -1 padding.
constexpr int EXP_THREADS = 128;
constexpr int EXP_MAX_N = 2048;

__global__ void __launch_bounds__(EXP_THREADS)
expand_kernel(const long long* __restrict__ code_rows, int Q, int n, const long long* __restrict__ csr_off,
              const long long* __restrict__ csr_rows, long long pitch, long long* __restrict__ cand_idx,
              long long* __restrict__ cand_off, long long* __restrict__ cand_cnt) {
  __shared__ long long s_start[EXP_MAX_N];   // first csr position of code j
  __shared__ long long s_dst[EXP_MAX_N + 1];  // exclusive prefix of the counts
  __shared__ long long s_carry;
  const int qi = blockIdx.x, tid = threadIdx.x;
  long long* seg = cand_idx + (long long)qi * pitch;
  if (tid == 0) {
    s_carry = 0;
    cand_off[qi] = (long long)qi * pitch;
    if (qi == Q - 1) cand_off[Q] = (long long)Q * pitch;
  }
  __syncthreads();
  // counts, then a chunked block scan (EXP_THREADS codes per pass)
  for (int base = 0; base < n; base += EXP_THREADS) {
    const int j = base + tid;
    long long cnt = 0, start = 0;
    if (j < n) {
      const long long c = code_rows[(long long)qi * n + j];
      if (c >= 0) {
        start = csr_off[c];
        cnt = csr_off[c + 1] - start;
      }
      s_start[j] = start;
    }
    // inclusive warp scan + cross-warp carry through shared memory
    long long inc = cnt;
    const int lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(sb::FULL_MASK, inc, o);
      if (lane >= o) inc += t;
    }
    __shared__ long long s_wsum[EXP_THREADS / 32];
    if (lane == 31) s_wsum[w] = inc;
    __syncthreads();
    long long woff = s_carry;
    for (int ww = 0; ww < w; ++ww) woff += s_wsum[ww];
    if (j < n) s_dst[j] = woff + inc - cnt;
    __syncthreads();
    if (tid == EXP_THREADS - 1) s_carry = woff + inc;
    __syncthreads();
  }
  const long long total = min(s_carry, pitch);
  if (tid == 0) {
    s_dst[n] = s_carry;
    cand_cnt[qi] = total;
  }
  __syncthreads();
  // copy: one warp per code, lanes over its rows
  const int lane = tid & 31, w = tid >> 5;
  for (int j = w; j < n; j += EXP_THREADS / 32) {
    const long long d0 = s_dst[j], cnt = s_dst[j + 1] - d0, src = s_start[j];
    for (long long i = lane; i < cnt; i += 32)
      if (d0 + i < pitch) seg[d0 + i] = csr_rows[src + i];
  }
  for (long long i = total + tid; i < pitch; i += EXP_THREADS) seg[i] = -1;
}

}  // namespace

extern "C" {

static int rerank_impl(const float* db, int64_t N, int32_t D, int64_t ldd, const float* q, int32_t Q, int64_t ldq,
                       const int64_t* cand_idx, const int64_t* cand_off, int64_t M, int32_t metric, double* out,
                       int64_t row_base, double miss_value, void* stream, const float* const* shards = nullptr,
                       const int64_t* shard_bounds = nullptr, int n_shards = 0, int64_t pitch = 0) {
  SB_REQUIRE(D >= 1 && Q >= 1 && M >= 0 && N >= 0, "sb_rerank: bad sizes");
  SB_REQUIRE(ldd >= D && ldq >= D, "sb_rerank: leading dimension smaller than D");
  SB_REQUIRE(metric >= SB_METRIC_EUCLIDEAN && metric <= SB_METRIC_HIK, "sb_rerank: bad metric %d", metric);
  if (M == 0) return SB_OK;
  SB_REQUIRE((db || shards) && q && cand_idx && cand_off && out, "sb_rerank: NULL pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // peer shards: the caller guarantees 16-byte aligned shard bases (checked in sb_rerank_peer)
  const int vec_ok = ((reinterpret_cast<uintptr_t>(db) | reinterpret_cast<uintptr_t>(q)) % 16 == 0 && ldd % 4 == 0 &&
                      ldq % 4 == 0) ? 1 : 0;
  const long long* sbnd = reinterpret_cast<const long long*>(shard_bounds);
  const unsigned grid = (unsigned)((M + RR_WARPS - 1) / RR_WARPS);
  const long long* ci = reinterpret_cast<const long long*>(cand_idx);
  const long long* co = reinterpret_cast<const long long*>(cand_off);
  sb::ProfScope prof("rerank_kernel", st);
  switch (metric) {
    case SB_METRIC_EUCLIDEAN:
      rerank_kernel<SB_METRIC_EUCLIDEAN><<<grid, RR_WARPS * 32, 0, st>>>(db, N, D, ldd, q, Q, ldq, ci, co, M, out, vec_ok,
                                                                         row_base, miss_value, shards, sbnd, n_shards, (long long)pitch);
      break;
    case SB_METRIC_COSINE:
      rerank_kernel<SB_METRIC_COSINE><<<grid, RR_WARPS * 32, 0, st>>>(db, N, D, ldd, q, Q, ldq, ci, co, M, out, vec_ok,
                                                                      row_base, miss_value, shards, sbnd, n_shards, (long long)pitch);
      break;
    default:
      rerank_kernel<SB_METRIC_HIK><<<grid, RR_WARPS * 32, 0, st>>>(db, N, D, ldd, q, Q, ldq, ci, co, M, out, vec_ok,
                                                                   row_base, miss_value, shards, sbnd, n_shards, (long long)pitch);
      break;
  }
  sb::count_launch();
  return sb::check_launch("rerank_kernel");
}

int sb_rerank(const float* db, int64_t N, int32_t D, int64_t ldd, const float* q, int32_t Q, int64_t ldq,
              const int64_t* cand_idx, const int64_t* cand_off, int64_t M, int32_t metric, double* out,
              void* stream) {
  return rerank_impl(db, N, D, ldd, q, Q, ldq, cand_idx, cand_off, M, metric, out, 0, nan(""), stream);
}

int sb_rerank_base(const float* db, int64_t N, int64_t row_base, int32_t D, int64_t ldd, const float* q, int32_t Q,
                   int64_t ldq, const int64_t* cand_idx, const int64_t* cand_off, int64_t M, int32_t metric,
                   double* out, void* stream) {
  return rerank_impl(db, N, D, ldd, q, Q, ldq, cand_idx, cand_off, M, metric, out, row_base, nan(""), stream);
}

int sb_rerank_shard(const float* db, int64_t N, int64_t row_base, int32_t D, int64_t ldd, const float* q, int32_t Q,
                    int64_t ldq, const int64_t* cand_idx, const int64_t* cand_off, int64_t M, int32_t metric,
                    double* out, void* stream) {
  return rerank_impl(db, N, D, ldd, q, Q, ldq, cand_idx, cand_off, M, metric, out, row_base, 0.0, stream);
}

/* Re-rank against a ROW-SHARDED table whose shards live on several GPUs of one box: shards[s] (device
 * array of n_shards pointers, valid in THIS process: the local shard and CUDA-IPC mappings of the peers')
 * holds global rows [shard_bounds[s], shard_bounds[s+1]).  Candidate rows are read where they live, over
 * NVLink for peer shards -- the re-rank needs no collective.  Rows outside every shard give NaN. */
int sb_rerank_peer(const float* const* shards, const int64_t* shard_bounds, int32_t n_shards, int32_t D, int64_t ldd,
                   const float* q, int32_t Q, int64_t ldq, const int64_t* cand_idx, const int64_t* cand_off, int64_t M,
                   int64_t pitch, int32_t metric, int32_t aligned16, double* out, void* stream) {
  SB_REQUIRE(shards && shard_bounds && n_shards >= 1, "sb_rerank_peer: NULL shard table");
  SB_REQUIRE(pitch == 0 || M == (int64_t)Q * pitch, "sb_rerank_peer: fixed-pitch layout needs M == Q * pitch");
  // `aligned16`: every shard base is 16-byte aligned (the host knows the pointers; the device array is not read here)
  const float* fake_db = aligned16 ? reinterpret_cast<const float*>(uintptr_t(16)) : reinterpret_cast<const float*>(uintptr_t(4));
  return rerank_impl(fake_db, 0, D, ldd, q, Q, ldq, cand_idx, cand_off, M, metric, out, 0, nan(""), stream, shards,
                     shard_bounds, n_shards, pitch);
}

/* sb_rerank for the FIXED-PITCH candidate layout of sb_expand_candidates (query of slot j = j / pitch: no
 * search for the owning query); padding slots (-1) give NaN. */
int sb_rerank_pitched(const float* db, int64_t N, int32_t D, int64_t ldd, const float* q, int32_t Q, int64_t ldq,
                      const int64_t* cand_idx, const int64_t* cand_off, int64_t pitch, int32_t metric, double* out,
                      void* stream) {
  SB_REQUIRE(pitch >= 1, "sb_rerank_pitched: pitch must be positive");
  return rerank_impl(db, N, D, ldd, q, Q, ldq, cand_idx, cand_off, (int64_t)Q * pitch, metric, out, 0, nan(""), stream,
                     nullptr, nullptr, 0, pitch);
}

/* Peer access from the CURRENT device to `peer_device` (needed before kernels read CUDA-IPC mappings of a
 * peer's memory).  Returns SB_OK when access is (already) enabled, SB_ERR_UNSUPPORTED when the devices
 * cannot reach each other. */
int sb_enable_peer_access(int32_t peer_device) {
  int cur = 0;
  SB_CUDA_TRY(cudaGetDevice(&cur));
  if (cur == peer_device) return SB_OK;
  int can = 0;
  SB_CUDA_TRY(cudaDeviceCanAccessPeer(&can, cur, peer_device));
  if (!can) {
    sb::set_error("sb_enable_peer_access: device %d cannot access device %d", cur, peer_device);
    return SB_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return SB_OK;
  }
  if (e != cudaSuccess) {
    sb::set_error("cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
    return SB_ERR_CUDA;
  }
  return SB_OK;
}

int sb_rerank_select(const double* dist, const int64_t* cand_off, int32_t Q, int32_t n, int64_t* out_pos,
                     double* out_dist, void* stream) {
  SB_REQUIRE(Q >= 1 && n >= 1, "sb_rerank_select: bad sizes");
  SB_REQUIRE(cand_off && out_pos && out_dist, "sb_rerank_select: NULL pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  sb::ProfScope prof("rerank_select_kernel", st);
  rerank_select_kernel<<<Q, SEL_THREADS, 0, st>>>(dist, reinterpret_cast<const long long*>(cand_off), nullptr, nullptr, Q,
                                                  n, reinterpret_cast<long long*>(out_pos), out_dist, 0);
  sb::count_launch();
  return sb::check_launch("rerank_select_kernel");
}

int sb_rerank_select_rows(const double* dist, const int64_t* cand_off, const int64_t* cand_cnt, const int64_t* cand_idx,
                          int32_t Q, int32_t n, int32_t tie_by_row, int64_t* out_rows, double* out_dist, void* stream) {
  SB_REQUIRE(Q >= 1 && n >= 1, "sb_rerank_select_rows: bad sizes");
  SB_REQUIRE(cand_off && cand_idx && out_rows && out_dist, "sb_rerank_select_rows: NULL pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  sb::ProfScope prof("rerank_select_kernel", st);
  rerank_select_kernel<<<Q, SEL_THREADS, 0, st>>>(dist, reinterpret_cast<const long long*>(cand_off),
                                                  reinterpret_cast<const long long*>(cand_cnt),
                                                  reinterpret_cast<const long long*>(cand_idx), Q, n,
                                                  reinterpret_cast<long long*>(out_rows), out_dist, tie_by_row ? 1 : 0);
  sb::count_launch();
  return sb::check_launch("rerank_select_kernel");
}

int sb_expand_candidates(const int64_t* code_rows, int32_t Q, int32_t n, const int64_t* csr_off, const int64_t* csr_rows,
                         int64_t pitch, int64_t* cand_idx, int64_t* cand_off, int64_t* cand_cnt, void* stream) {
  SB_REQUIRE(Q >= 1 && n >= 1 && pitch >= 1, "sb_expand_candidates: bad sizes");
  SB_REQUIRE(n <= EXP_MAX_N, "sb_expand_candidates: n=%d exceeds the supported maximum of %d", n, EXP_MAX_N);
  SB_REQUIRE(code_rows && csr_off && csr_rows && cand_idx && cand_off && cand_cnt, "sb_expand_candidates: NULL pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  sb::ProfScope prof("expand_kernel", st);
  expand_kernel<<<Q, EXP_THREADS, 0, st>>>(reinterpret_cast<const long long*>(code_rows), Q, n,
                                           reinterpret_cast<const long long*>(csr_off),
                                           reinterpret_cast<const long long*>(csr_rows), pitch,
                                           reinterpret_cast<long long*>(cand_idx), reinterpret_cast<long long*>(cand_off),
                                           reinterpret_cast<long long*>(cand_cnt));
  sb::count_launch();
  return sb::check_launch("expand_kernel");
}

}  // extern "C"
