// ITQ training on the tensor cores: G[b][D] = UX^T . (X / div - mean), the row-count-scaled
// product of every ITQ iteration (reference: smqtk_indexing/impls/lsh_functor/itq.py:274,
// c = ux^T . v; in the V-free form of fit.py::itq_rotation_streaming c = (ux^T x) . pc_top).
// UX is the +-1 matrix of the current sign bits (packed codes), X the float32 training matrix.
//
// The contraction runs over ROWS, so both operands are "MN-major" for the tensor core: for a
// fixed row k the neighbouring bits / features are contiguous.  For 32-bit MN-major operands the
// only shared-memory layout the tensor core accepts is SWIZZLE_128B_BASE32B (descriptor layout
// type 1): blocks of 32 MN elements (128 bytes) x 4 rows (k) = 512-byte atoms, the 32-byte chunk
// index XOR-ed with (k % 4); MN blocks LBO apart, 4-row k groups SBO apart.  A warp stores one
// row of 32 neighbouring 16-byte groups = 4 x 128 permuted-contiguous bytes: bank-conflict free.
//   A = UX^T : M = 128 bits per sub-tile (b <= 256 -> 1 or 2 sub-tiles), +-1.0 exact in TF32
//   B = X    : N = D tile (256 features for one sub-tile, 128 for two), a = x/div - mean formed in
//              FP32, scaled by a power of two so that |a| <= 1, and split on FIXED-POINT grids:
//              hi = multiple of 2^-10 (|k| <= 2^10: exact in TF32), lo = a - hi rounded to a multiple
//              of 2^-21 (|k| <= 2^10 again).
//   D        : FP32 in TMEM, SEPARATE accumulators for the hi and the lo products.  The tensor core
//              truncates when it aligns addends to the accumulator, which for a plain hi + lo
//              accumulation shrinks the large entries of G by ~1e-5 -- enough to move the polar
//              factor of G.P by 1e-4 per ITQ iteration.  On a common grid every addend is an
//              integer multiple of the accumulator's last bit as long as |sum| < 2^24 grid units,
//              i.e. for windows of FLUSH_ROWS <= 2^13 rows: the accumulation is EXACT, and the only
//              error left is the 2^-22 (relative to the bound) rounding of a.  Each window is
//              flushed into an FP64 partial (one per CTA); partials are reduced in a fixed order
//              afterwards (deterministic).
// One persistent CTA per (row range, D tile); warps: 0-3 flush epilogue, 4 MMA issue + TMEM,
// 6-21 producers (coalesced 128-bit loads of X and of the codes, transform, STS.128).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

using namespace tcptx;

namespace {

constexpr int KS = 16;                      // rows (K) per pipeline stage = two MMA K steps of 8
constexpr int MN_BLOCK = KS * 128;          // one block of 32 MN elements, all 16 rows of the stage: 2048 B (= LBO)
constexpr int K_GROUP = 512;                // 4 rows x 128 B (= SBO)
constexpr int A_SUB = 4 * MN_BLOCK;         // one 128-bit sub-tile, 16 rows   -> 8192 B
constexpr int EPI_WARPS = 4, MMA_WARP = 4, XF_WARP0 = 6, XF_WARPS = 16;
constexpr int THREADS = (XF_WARP0 + XF_WARPS) * 32;   // 704
constexpr int MAX_STAGES = 6;
constexpr int FLUSH_ROWS = 8192;             // x 2^10 grid units per row < 2^24: exact FP32 accumulation

struct GramParams {
  const uint32_t* codes;   // u32[n][W]
  int W, b, subs;          // subs = ceil(b / 128)
  const float* X;          // f32[n][ldx]
  long long ldx, n;
  int D, Dt;               // Dt = columns of this launch's D tiles (<= 256, multiple of 16)
  const float* mean;       // f32[D] or NULL
  const float* row_div;    // f32[n] or NULL
  double* partial;         // f64[row_chunks][b][D]
  long long rows_per_cta;  // multiple of KS
  int stages;
  int tmem_cols;
  int tile_d;              // features per CTA: 256 (one sub-tile) or 128 (two)
  float inv_scale;         // power of two, |x/div - mean| * inv_scale <= 1
  double scale;
};

// MN-major SWIZZLE_128B_BASE32B shared-memory matrix descriptor (LBO = MN block stride, SBO = k group stride)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((MN_BLOCK >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((K_GROUP >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;                             // descriptor version (sm_100)
  d |= 1ull << 61;                             // layout_type = SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of the 16-byte group g (4 MN elements) of row kr inside a stage region
__device__ __forceinline__ uint32_t mn_offset(int g, int kr) {
  const uint32_t chunk32 = (uint32_t)((g >> 1) & 3) ^ (uint32_t)(kr & 3);
  return (uint32_t)(g >> 3) * MN_BLOCK + (uint32_t)(kr >> 2) * K_GROUP + (uint32_t)(kr & 3) * 128 + chunk32 * 32 +
         (uint32_t)(g & 1) * 16;
}

// v (|v| <= 1) -> hi on the 2^-10 grid, lo = v - hi on the 2^-21 grid (round to nearest through the
// magic-number add; v - hi is exact in FP32).  Both have at most 11 significant bits: exact TF32.
__device__ __forceinline__ void split_grid(float v, uint32_t& hi, uint32_t& lo) {
  const float MH = 12288.0f;                   // 1.5 * 2^13: FP32 last bit = 2^-10
  const float ML = 6.0f;                       // 1.5 * 2^2 : FP32 last bit = 2^-21
  const float h = __fsub_rn(__fadd_rn(v, MH), MH);
  const float l = __fsub_rn(__fadd_rn(__fsub_rn(v, h), ML), ML);
  hi = __float_as_uint(h);
  lo = __float_as_uint(l);
}

__global__ void __launch_bounds__(THREADS, 1) gram_bits_tc_kernel(const GramParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms: align by hand
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d0 = blockIdx.y * p.tile_d;                    // first feature of this CTA's D tile
  const int Dt = min(p.tile_d, p.D - d0);
  const int ngroups = Dt / 4;                              // 16-byte groups of B per row
  const uint32_t b_part = (uint32_t)(Dt / 32) * MN_BLOCK;  // B_hi (or B_lo) per stage
  const uint32_t a_bytes = (uint32_t)p.subs * A_SUB;
  const uint32_t stage_bytes = a_bytes + 2 * b_part;       // [A sub0 | A sub1 | B_hi | B_lo]
  unsigned char* bars_raw = smem + (size_t)p.stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bars_raw);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + MAX_STAGES * 8;
  const uint32_t acc_full = empty0 + MAX_STAGES * 8, acc_empty = acc_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_raw + (2 * MAX_STAGES + 2) * 8);
  const uint32_t smem0 = smem_u32(smem);

  const long long r0 = (long long)blockIdx.x * p.rows_per_cta;
  const long long r1 = min(p.n, r0 + p.rows_per_cta);
  const long long nstages = (r1 > r0) ? (r1 - r0 + KS - 1) / KS : 0;
  const int flush_stages = FLUSH_ROWS / KS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full0 + s * 8, XF_WARPS); mbar_init(empty0 + s * 8, 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= XF_WARP0) {
    // =========================== producers ===========================
    // thread -> (row kr = xt / 32 of the stage, lane c): X float4 groups c and c + 32 of that row,
    // code bit groups c and c + 32 (4 bits each) of that row.
    const int xt = threadIdx.x - XF_WARP0 * 32;
    const int kr = xt >> 5, c = xt & 31;
    int stage = 0;
    uint32_t phase = 0;
    // prefetch of the next stage's X values and code words
    float4 xv[2], xn_[2];
    uint32_t cw[2], cn_[2];
    auto load = [&](long long st, float4 (&x)[2], uint32_t (&w)[2]) {
      const long long row = r0 + st * KS + kr;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int g = c + 32 * j;                            // MN group index
        x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        w[j] = 0u;
        if (row < r1) {
          if (g < ngroups) {
            float4 v = ldg_stream(p.X + row * p.ldx + d0 + g * 4);
            if (p.row_div) {
              const float dv = __ldg(p.row_div + row);
              v.x = __fdiv_rn(v.x, dv); v.y = __fdiv_rn(v.y, dv); v.z = __fdiv_rn(v.z, dv); v.w = __fdiv_rn(v.w, dv);
            }
            if (p.mean) {
              const float4 m4 = __ldg(reinterpret_cast<const float4*>(p.mean + d0 + g * 4));
              v.x -= m4.x; v.y -= m4.y; v.z -= m4.z; v.w -= m4.w;
            }
            x[j] = v;
          }
          // bits m = 4g .. 4g+3 (vector index; integer bit position b-1-m), m < b
          const int m = 4 * g;
          if (m < p.b) {
            const int pbit = p.b - 1 - m;                    // position of bit m; m+1..m+3 are pbit-1..pbit-3
            const uint32_t word = __ldg(p.codes + row * p.W + (p.W - 1 - pbit / 32));
            w[j] = (word >> ((pbit & 31) - 3)) & 0xFu;       // b % 4 == 0 -> the 4 bits share a word; bit 3 = m
          }
        }
      }
    };
    if (nstages > 0) load(0, xv, cw);
    for (long long st = 0; st < nstages; ++st) {
      if (st + 1 < nstages) load(st + 1, xn_, cn_);
      mbar_wait(empty0 + stage * 8, phase ^ 1);
      unsigned char* sa = smem + (size_t)stage * stage_bytes;
      unsigned char* sb_hi = sa + a_bytes;
      const long long row = r0 + st * KS + kr;
      const bool rvalid = row < r1;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int g = c + 32 * j;
        if (g < ngroups) {
          const float4 v = xv[j];
          uint4 hi, lo;
          split_grid(v.x * p.inv_scale, hi.x, lo.x);
          split_grid(v.y * p.inv_scale, hi.y, lo.y);
          split_grid(v.z * p.inv_scale, hi.z, lo.z);
          split_grid(v.w * p.inv_scale, hi.w, lo.w);
          const uint32_t off = mn_offset(g, kr);
          *reinterpret_cast<uint4*>(sb_hi + off) = hi;
          *reinterpret_cast<uint4*>(sb_hi + b_part + off) = lo;
        }
        if (g < 32 * p.subs) {                               // bit group g: sub-tile g / 32, MN group g % 32
          const uint32_t pos = __float_as_uint(1.0f), neg = __float_as_uint(-1.0f);
          const uint32_t w4 = cw[j];
          uint4 a;
          const bool live = rvalid && (4 * g < p.b);
          a.x = live ? ((w4 & 8u) ? pos : neg) : 0u;          // bit m
          a.y = live ? ((w4 & 4u) ? pos : neg) : 0u;          // bit m+1
          a.z = live ? ((w4 & 2u) ? pos : neg) : 0u;
          a.w = live ? ((w4 & 1u) ? pos : neg) : 0u;
          *reinterpret_cast<uint4*>(sa + (uint32_t)(g >> 5) * A_SUB + mn_offset(g & 31, kr)) = a;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full0 + stage * 8);
      xv[0] = xn_[0]; xv[1] = xn_[1]; cw[0] = cn_[0]; cw[1] = cn_[1];
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issue ===========================
    if (lane == 0) {
      // D = F32, A = B = TF32, both MN-major (bits 15, 16), N = Dt, M = 128
      uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(Dt >> 3) << 17) |
                       ((uint32_t)(128 >> 4) << 24);
      const uint32_t lo_col = (uint32_t)(p.subs * p.tile_d);   // TMEM: [hi sub0 | hi sub1 | lo sub0 | lo sub1]
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (long long st = 0; st < nstages; ++st) {
        const bool first = (st % flush_stages) == 0;
        if (first) {
          mbar_wait(acc_empty, acc_phase ^ 1);               // the previous window has been flushed
          tc_fence_after();
        }
        mbar_wait(full0 + stage * 8, phase);
        tc_fence_after();
        const uint32_t sa = smem0 + stage * stage_bytes, sbh = sa + a_bytes, sbl = sbh + b_part;
#pragma unroll
        for (int kg = 0; kg < 2; ++kg) {                     // MMA K step = 8 rows = two 4-row k groups
          const uint64_t dbh = umma_desc_mn(sbh + kg * 2 * K_GROUP);
          const uint64_t dbl = umma_desc_mn(sbl + kg * 2 * K_GROUP);
          for (int sub = 0; sub < p.subs; ++sub) {
            const uint64_t da = umma_desc_mn(sa + sub * A_SUB + kg * 2 * K_GROUP);
            const uint32_t d_tmem = tmem_base + (uint32_t)(sub * p.tile_d);
            const uint32_t acc = (first && kg == 0) ? 0u : 1u;
            umma_tf32(d_tmem, da, dbh, idesc, acc);
            umma_tf32(d_tmem + lo_col, da, dbl, idesc, acc);
          }
        }
        umma_commit(empty0 + stage * 8);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if ((st + 1) % flush_stages == 0 || st + 1 == nstages) {
          umma_commit(acc_full);
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp < EPI_WARPS) {
    // =========================== flush: TMEM (FP32) -> FP64 partial of this CTA ===========================
    uint32_t acc_phase = 0;
    const long long windows = (nstages + flush_stages - 1) / flush_stages;
    double* part = p.partial + (size_t)blockIdx.x * p.b * p.D;
    for (long long wdw = 0; wdw < windows; ++wdw) {
      mbar_wait(acc_full, acc_phase);
      tc_fence_after();
      for (int sub = 0; sub < p.subs; ++sub) {
        const int m = sub * 128 + warp * 32 + lane;          // bit index = TMEM lane
#pragma unroll 1
        for (int c16 = 0; c16 < Dt / 16; ++c16) {
          uint32_t v[16], w[16];
          const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(sub * p.tile_d + c16 * 16);
          tmem_ld16_nowait(taddr, v);
          tmem_ld16_nowait(taddr + (uint32_t)(p.subs * p.tile_d), w);
          tmem_ld_wait();
          if (m < p.b) {
            double* dst = part + (size_t)m * p.D + d0 + c16 * 16;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              dst[j] += p.scale * ((double)__uint_as_float(v[j]) + (double)__uint_as_float(w[j]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

__global__ void gram_reduce_kernel(const double* __restrict__ partial, int parts, long long elems, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= elems) return;
  double acc = 0.0;
  for (int q = 0; q < parts; ++q) acc += partial[(size_t)q * elems + i];
  out[i] = acc;
}

struct GramPlan {
  int d_tiles, row_chunks, stages, subs, tile_d;
  long long rows_per_cta;
  size_t smem_bytes, workspace_bytes;
};

GramPlan make_gram_plan(int64_t n, int32_t D, int32_t b) {
  GramPlan g;
  g.subs = (b + 127) / 128;
  g.tile_d = g.subs == 1 ? 256 : 128;                        // hi + lo accumulators: 2 * subs * tile_d = 512 TMEM columns
  g.d_tiles = (D + g.tile_d - 1) / g.tile_d;
  int chunks = sb::sm_count() / g.d_tiles;
  if (chunks < 1) chunks = 1;
  long long rpc = (n + chunks - 1) / chunks;
  rpc = (rpc + FLUSH_ROWS - 1) / FLUSH_ROWS * FLUSH_ROWS;    // whole flush windows per CTA
  g.rows_per_cta = rpc;
  g.row_chunks = (int)((n + rpc - 1) / rpc);
  const int Dt = D < g.tile_d ? D : g.tile_d;
  const size_t stage = (size_t)g.subs * A_SUB + 2 * (size_t)(Dt / 32) * MN_BLOCK;
  const size_t tail = 1024 + (2 * MAX_STAGES + 2) * 8 + 16;
  int s = (int)((220 * 1024 - tail) / stage);
  if (s > MAX_STAGES) s = MAX_STAGES;
  g.stages = s;
  g.smem_bytes = (size_t)s * stage + tail;
  g.workspace_bytes = (size_t)g.row_chunks * b * D * sizeof(double);
  return g;
}

}  // namespace

extern "C" {

int sb_fit_gram_bits_tc_supported(int64_t n, int32_t D, int64_t ldx, int32_t b) {
  return n >= 1 && D >= 32 && D % 32 == 0 && ldx % 4 == 0 && b >= 4 && b % 4 == 0 && b <= 256;
}

size_t sb_fit_gram_bits_tc_workspace_bytes(int64_t n, int32_t D, int32_t b) {
  if (n < 1 || D < 32 || b < 4 || b > 256) return 0;
  return make_gram_plan(n, D, b).workspace_bytes;
}

// out f64[b][D] = sum_r (bit_m(codes[r]) ? +1 : -1) * (X[r][d] / row_div[r] - mean[d])
int sb_fit_gram_bits_tc(const uint32_t* codes, int32_t W, int32_t b, const float* X, int64_t n, int32_t D, int64_t ldx,
                        const float* mean, const float* row_div, float bound, double* out, void* workspace,
                        size_t workspace_bytes, void* stream) {
  SB_REQUIRE(codes && X && out, "sb_fit_gram_bits_tc: NULL pointer");
  SB_REQUIRE(W * 32 >= b, "sb_fit_gram_bits_tc: %d words cannot hold %d bits", W, b);
  if (!sb_fit_gram_bits_tc_supported(n, D, ldx, b) || (reinterpret_cast<uintptr_t>(X) & 15u) ||
      (mean && (reinterpret_cast<uintptr_t>(mean) & 15u))) {
    sb::set_error("sb_fit_gram_bits_tc: needs D %% 32 == 0, ldx %% 4 == 0, b %% 4 == 0, b <= 256, 16-byte aligned X / mean");
    return SB_ERR_UNSUPPORTED;
  }
  SB_REQUIRE(bound > 0.f && bound < 1e30f, "sb_fit_gram_bits_tc: bound must be a finite positive bound on |x / div - mean|");
  const GramPlan g = make_gram_plan(n, D, b);
  if (workspace == nullptr || workspace_bytes < g.workspace_bytes) {
    sb::set_error("sb_fit_gram_bits_tc: workspace too small (%zu < %zu bytes)", workspace_bytes, g.workspace_bytes);
    return SB_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SB_CUDA_TRY(cudaMemsetAsync(workspace, 0, g.workspace_bytes, st));
  GramParams p;
  p.codes = codes; p.W = W; p.b = b; p.subs = g.subs; p.X = X; p.ldx = ldx; p.n = n; p.D = D; p.Dt = D < g.tile_d ? D : g.tile_d;
  p.mean = mean; p.row_div = row_div; p.partial = static_cast<double*>(workspace); p.rows_per_cta = g.rows_per_cta;
  p.stages = g.stages;
  p.tmem_cols = 512;
  p.tile_d = g.tile_d;
  {
    int e = 0;
    frexpf(bound, &e);                                       // bound = f * 2^e, f in [0.5, 1)  ->  bound <= 2^e
    p.inv_scale = ldexpf(1.0f, -e);
    p.scale = ldexp(1.0, e);
  }
  SB_CUDA_TRY(cudaFuncSetAttribute(gram_bits_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
  {
    sb::ProfScope prof("gram_bits_tc_kernel", st);
    dim3 grid(g.row_chunks, g.d_tiles);
    gram_bits_tc_kernel<<<grid, THREADS, g.smem_bytes, st>>>(p);
    sb::count_launch();
    if (int rc = sb::check_launch("gram_bits_tc_kernel")) return rc;
  }
  const long long elems = (long long)b * D;
  gram_reduce_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(p.partial, g.row_chunks, elems, out);
  sb::count_launch();
  return sb::check_launch("gram_reduce_kernel");
}

}  // extern "C"
