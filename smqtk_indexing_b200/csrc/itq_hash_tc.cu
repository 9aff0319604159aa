// Stage 1 of the LSH path on the 5th-generation tensor cores:
//   codes[r] = pack( ((X[r] / div[r]) - mean) . R  >= 0 )
// Replaces ItqFunctor.get_hash (reference: smqtk_indexing/impls/lsh_functor/
// itq.py:389-408) + bit_vector_to_int_large (smqtk_indexing/utils/bits.py:4-20)
// for whole matrices of descriptors -- the only dense contraction on the path.
//
// Arithmetic: 3xTF32.  a = x/div - mean is formed in FP32 (the reference subtracts
// in the input dtype, itq.py:404), then split a = a_hi + a_lo with a_hi = tf32(a),
// a_lo = tf32(a - a_hi); R is split the same way once (sb_itq_rotation_image).
// z = a_hi.r_hi + a_lo.r_hi + a_hi.r_lo accumulated in FP32 in tensor memory: the
// dropped a_lo.r_lo term and the rounding of the lo parts are ~2^-22 relative,
// i.e. FP32-class accuracy on the TF32 tensor pipe (hard part 3).
//
// Mapping (sm_100a, one persistent CTA per SM, 22 warps):
//   tile      = 256 descriptor rows (two UMMA M=128 sub-tiles) x all b columns;
//               accumulators: 2 x b FP32 columns of TMEM.
//   stage     = 16 K-elements: A_hi, A_lo (256 x 16 tf32 each) + R_hi, R_lo
//               (b x 16 each) in shared memory, canonical K-major no-swizzle UMMA
//               layout (8-row x 16-byte core matrices).  Both sub-tiles reuse the
//               stage's R operand, which halves the L2 -> SMEM traffic for R.
//   warps 0-3   epilogue: tcgen05.ld 32 lanes x 32 columns -> sign test -> one
//               32-bit word per thread per 32 columns -> coalesced 128-bit stores.
//   warp  4     TMEM allocation + single-thread tcgen05.mma issue (12 per stage),
//               tcgen05.commit to the stage's "empty" mbarrier / "accumulator full".
//   warp  5     R operand: one cp.async.bulk (TMA) per stage from the pre-split image.
//   warps 6-21  A operand: 128-bit streaming LDG of raw X (4 stages of register
//               prefetch), normalise/centre/split in registers, conflict-free STS
//               into the UMMA layout, fence.proxy.async, mbarrier arrive.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace sb {
int itq_hash_tc_supported(int64_t n, int32_t D, int64_t ldx, int32_t b, const float* X);
}

namespace {

constexpr int KC = 16;                 // K elements per pipeline stage
constexpr int KCHUNKS = KC / 4;        // 16-byte chunks along K per stage
constexpr int TILE_M = 256;            // rows per CTA tile (2 x UMMA_M)
constexpr int UMMA_M = 128;
constexpr int UMMA_K = 8;              // tf32: 32 bytes of K per instruction
constexpr int MAX_STAGES = 8;
constexpr int EPI_WARPS = 4;
constexpr int MMA_WARP = 4;
constexpr int LOAD_WARP = 5;
constexpr int XF_WARP0 = 6;
constexpr int XF_WARPS = 16;
constexpr int THREADS = (XF_WARP0 + XF_WARPS) * 32;   // 704
constexpr int XF_ROWGROUPS = (TILE_M / 8) / XF_WARPS;  // 8-row groups per producer warp and stage = 2
constexpr int PF = 4;                  // register prefetch depth of the A path (stages)
constexpr int A_PLANE = TILE_M * 16;   // bytes of one 16-byte K-chunk plane of A (all 256 rows)
constexpr int A_PART = KCHUNKS * A_PLANE;             // A_hi (or A_lo) bytes per stage = 16 KB

using namespace tcptx;

// ---- one-off: split R into the per-stage shared-memory image ------------------
// image[kc][part][chunk][n][4] (part 0 = hi, 1 = lo): exactly the bytes a stage's R
// region holds, so the main kernel fetches a stage with ONE bulk copy.
__global__ void rotation_image_kernel(const float* __restrict__ R, int D, int b, uint32_t* __restrict__ img) {
  const long long total = (long long)D * b;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i / b), n = (int)(i - (long long)k * b);
    const float r = R[i];
    const uint32_t hi = to_tf32(r);
    const uint32_t lo = to_tf32(r - __uint_as_float(hi));
    const int kc = k / KC, c = (k % KC) / 4, e = k & 3;
    const size_t base = ((size_t)kc * 2) * KCHUNKS * b * 4;   // in 32-bit words
    const size_t off = ((size_t)c * b + n) * 4 + e;
    img[base + off] = hi;
    img[base + (size_t)KCHUNKS * b * 4 + off] = lo;
  }
}

struct TcParams {
  const float* X;
  long long n;
  int D;
  long long ldx;
  const float* mean;      // may be NULL
  const float* row_div;   // may be NULL: x is used as is
  const uint32_t* r_image;
  int b;
  uint32_t* codes;
  int W;
  float* z_out;           // may be NULL
  int stages;
  int tmem_cols;
  // column blocks: a tile is (column block jb, 256-row tile rt); block jb uses the image at
  // r_image + jb * image_block_words.  Hashing has one block.
  int col_blocks;
  long long image_block_words;
  // EPI_L2 only: approximate squared-L2 filter (flat_l2.cu)
  const float* xn;        // f32[n]  |x|^2 per row
  const float* tq;        // f32[col_blocks * 256]  pass threshold per query (tau + margin) baked into the image
  unsigned long long* cand_buf;   // u64[queries][cap]: (float bits of d2) << 32 | row
  int* cand_cnt;          // i32[queries]
  int cap;
  unsigned int row_base;  // added to the row stored in the key
  int passes;             // 3 = 3xTF32 everywhere; 1 = single TF32 pass on the data chunks (EPI_L2: the
                          // synthetic threshold chunk always gets the full split)
};

constexpr int EPI_HASH = 0;   // sign test + bit-pack (ItqFunctor.get_hash)
constexpr int EPI_L2 = 1;     // d2 = |x|^2 + |q|^2 - 2 x.q against a per-query threshold (flat L2 index)

template <bool HAS_DIV, int EPI>
__global__ void __launch_bounds__(THREADS, 1) itq_hash_tc_kernel(const TcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = p.b;
  const uint32_t b_part = (uint32_t)KCHUNKS * b * 16;          // R_hi (or R_lo) bytes per stage
  const uint32_t stage_bytes = 2 * A_PART + 2 * b_part;
  // layout: [stages x (A_hi | A_lo | R_hi | R_lo)] [barriers] [tmem base]
  unsigned char* bars_raw = smem + (size_t)p.stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bars_raw);
  const uint32_t full0 = smem_u32(bars);                       // full[s]  : A written + R landed
  const uint32_t empty0 = full0 + MAX_STAGES * 8;              // empty[s] : MMAs reading stage s retired
  const uint32_t acc_full = empty0 + MAX_STAGES * 8;           // accumulators complete
  const uint32_t acc_empty = acc_full + 8;                     // epilogue drained TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_raw + (2 * MAX_STAGES + 2) * 8);
  const uint32_t smem0 = smem_u32(smem);

  const long long row_tiles = (p.n + TILE_M - 1) / TILE_M;
  const long long tiles = row_tiles * p.col_blocks;            // flat index t = jb * row_tiles + rt
  // EPI_L2 appends one synthetic K chunk: A = (-|x|^2/2, 1, 0...), B = (1, -h[query], 0...), so that
  // the accumulator is x.q - |x|^2/2 - h and the threshold test is a sign test
  const int nk = p.D / KC + (EPI == EPI_L2 ? 1 : 0);
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const long long my_tiles = (tiles > (long long)blockIdx.x) ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long items = my_tiles * nk;                       // (tile, k-chunk) pipeline items

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + s * 8, XF_WARPS + 1);
      mbar_init(empty0 + s * 8, 1);
    }
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
    // =========================== A operand producers ===========================
    const int xw = warp - XF_WARP0;
    const int r8 = lane & 7, c = lane >> 3;                     // row within an 8-row group, 16-byte K chunk
    // This thread's rows within a tile: (xw * XF_ROWGROUPS + i) * 8 + r8.  The load
    // cursor runs PF items ahead of the store cursor; both walk (tile, k-chunk).
    float4 buf[PF][XF_ROWGROUPS];
    float4 mbuf[PF];
    float dbuf[PF][XF_ROWGROUPS];
    long long ld_tile = blockIdx.x % row_tiles;                 // load cursor: ROW tile of flat tile blockIdx.x
    const long long ld_step = gridDim.x % row_tiles;
    int ld_kc = 0;
    long long ld_left = items;
    auto load_next = [&](float4 (&dst)[XF_ROWGROUPS], float4& m4, float (&dv)[XF_ROWGROUPS]) {
      if (ld_left <= 0) return;
      --ld_left;
      const long long row0 = ld_tile * TILE_M + xw * (XF_ROWGROUPS * 8) + r8;
      const float* src = p.X + row0 * p.ldx + ld_kc * KC + c * 4;
#pragma unroll
      for (int i = 0; i < XF_ROWGROUPS; ++i) {
        const long long row = row0 + i * 8;
        if (EPI == EPI_L2 && ld_kc == nk - 1) {
          dst[i] = (row < p.n && c == 0) ? make_float4(-0.5f * __ldg(p.xn + row), 1.0f, 0.f, 0.f)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          dst[i] = (row < p.n) ? ldg_stream(src + (long long)i * 8 * p.ldx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (HAS_DIV) dv[i] = (row < p.n) ? __ldg(p.row_div + row) : 1.0f;
      }
      m4 = p.mean ? __ldg(reinterpret_cast<const float4*>(p.mean + ld_kc * KC + c * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (++ld_kc == nk) {
        ld_kc = 0;
        ld_tile += ld_step;
        if (ld_tile >= row_tiles) ld_tile -= row_tiles;
      }
    };
#pragma unroll
    for (int q = 0; q < PF; ++q) load_next(buf[q], mbuf[q], dbuf[q]);
    int stage = 0, st_kc = 0;
    uint32_t phase = 0;
    const uint32_t st_off = (uint32_t)c * A_PLANE + (uint32_t)(xw * (XF_ROWGROUPS * 8) + r8) * 16;
    for (long long base = 0; base < items; base += PF) {
#pragma unroll
      for (int q = 0; q < PF; ++q) {
        if (base + q < items) {
          const float4 m4 = mbuf[q];
          const bool need_lo = (p.passes == 3) || (EPI == EPI_L2 && st_kc == nk - 1);
          if (++st_kc == nk) st_kc = 0;
          mbar_wait(empty0 + stage * 8, phase ^ 1);
          unsigned char* a_hi = smem + (size_t)stage * stage_bytes + st_off;
#pragma unroll
          for (int i = 0; i < XF_ROWGROUPS; ++i) {
            float4 v = buf[q][i];
            if (HAS_DIV) {
              const float dv = dbuf[q][i];
              v.x = __fdiv_rn(v.x, dv); v.y = __fdiv_rn(v.y, dv); v.z = __fdiv_rn(v.z, dv); v.w = __fdiv_rn(v.w, dv);
            }
            // rows past n carry (0 - mean): harmless, their results are never stored
            const float a0 = v.x - m4.x, a1 = v.y - m4.y, a2 = v.z - m4.z, a3 = v.w - m4.w;
            uint4 hi, lo;
            hi.x = to_tf32_finite(a0); hi.y = to_tf32_finite(a1); hi.z = to_tf32_finite(a2); hi.w = to_tf32_finite(a3);
            *reinterpret_cast<uint4*>(a_hi + i * 128) = hi;
            if (need_lo) {
              lo.x = to_tf32_finite(a0 - __uint_as_float(hi.x));
              lo.y = to_tf32_finite(a1 - __uint_as_float(hi.y));
              lo.z = to_tf32_finite(a2 - __uint_as_float(hi.z));
              lo.w = to_tf32_finite(a3 - __uint_as_float(hi.w));
              *reinterpret_cast<uint4*>(a_hi + A_PART + i * 128) = lo;
            }
          }
          fence_proxy_async();                                  // generic-proxy stores -> visible to the tensor core
          __syncwarp();
          if (lane == 0) mbar_arrive(full0 + stage * 8);
          load_next(buf[q], mbuf[q], dbuf[q]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == LOAD_WARP) {
    // =========================== R operand (TMA bulk) ===========================
    if (lane == 0) {
      int stage = 0, kc = 0;
      uint32_t phase = 0;
      long long t = blockIdx.x;
      const uint32_t* img = p.r_image + (t / row_tiles) * p.image_block_words;
      for (long long it = 0; it < items; ++it) {
        mbar_wait(empty0 + stage * 8, phase ^ 1);
        // hi part only when the chunk runs a single TF32 pass
        const uint32_t bytes = ((p.passes == 3) || (EPI == EPI_L2 && kc == nk - 1)) ? 2 * b_part : b_part;
        mbar_expect_tx(full0 + stage * 8, bytes);
        tma_bulk_g2s(smem0 + stage * stage_bytes + 2 * A_PART,
                     reinterpret_cast<const unsigned char*>(img) + (size_t)kc * 2 * b_part, bytes, full0 + stage * 8);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if (++kc == nk) {
          kc = 0;
          t += gridDim.x;
          img = p.r_image + (t / row_tiles) * p.image_block_words;
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issue (one thread) ===========================
    if (lane == 0) {
      // instruction descriptor: D=F32, A=B=TF32, both K-major, N = b, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(b >> 3) << 17) | ((uint32_t)(UMMA_M >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (long long t = 0; t < my_tiles; ++t) {
        mbar_wait(acc_empty, acc_phase ^ 1);                    // epilogue has drained the previous tile
        tc_fence_after();
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(full0 + stage * 8, phase);
          tc_fence_after();
          const uint32_t a_hi = smem0 + stage * stage_bytes, a_lo = a_hi + A_PART;
          const uint32_t r_hi = a_hi + 2 * A_PART, r_lo = r_hi + b_part;
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const uint32_t d_tmem = tmem_base + (uint32_t)(sub * b);
#pragma unroll
            for (int ks = 0; ks < KC / UMMA_K; ++ks) {
              const uint32_t a_off = (uint32_t)sub * (UMMA_M * 16) + (uint32_t)ks * 2 * A_PLANE;
              const uint32_t b_off = (uint32_t)ks * 2 * (uint32_t)b * 16;
              const uint64_t dah = umma_desc(a_hi + a_off, A_PLANE, 128);
              const uint64_t dal = umma_desc(a_lo + a_off, A_PLANE, 128);
              const uint64_t dbh = umma_desc(r_hi + b_off, (uint32_t)b * 16, 128);
              const uint64_t dbl = umma_desc(r_lo + b_off, (uint32_t)b * 16, 128);
              if ((p.passes == 3) || (EPI == EPI_L2 && kc == nk - 1)) {
                umma_tf32(d_tmem, dal, dbh, idesc, (kc | ks) ? 1u : 0u);   // small terms first
                umma_tf32(d_tmem, dah, dbl, idesc, 1u);
                umma_tf32(d_tmem, dah, dbh, idesc, 1u);
              } else {
                umma_tf32(d_tmem, dah, dbh, idesc, (kc | ks) ? 1u : 0u);
              }
            }
          }
          umma_commit(empty0 + stage * 8);                      // frees the stage when these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(acc_full);
        acc_phase ^= 1;
      }
    }
  } else {
    // =========================== epilogue (warps 0-3 = TMEM lane quadrants) ===========================
    const int nb = b >> 5;                                      // 32-column groups
    uint32_t acc_phase = 0;
    if constexpr (EPI == EPI_HASH) {
      const int wpad = p.W - nb;                                // leading zero words when W > b/32
      for (long long t = 0; t < my_tiles; ++t) {
        const long long tile = blockIdx.x + t * gridDim.x;      // one column block: flat tile == row tile
        mbar_wait(acc_full, acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const long long row = tile * TILE_M + sub * UMMA_M + warp * 32 + lane;
          uint32_t words[8];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            words[g] = 0u;
            if (g < nb) {
              uint32_t w = 0u;
#pragma unroll
              for (int h = 0; h < 2; ++h) {                     // 16 accumulator columns per TMEM load
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(sub * b + g * 32 + h * 16), v);
#pragma unroll
                for (int j = 0; j < 16; ++j) w |= (__uint_as_float(v[j]) >= 0.0f) ? (0x80000000u >> (h * 16 + j)) : 0u;
                if (p.z_out && row < p.n) {
                  float4* zo = reinterpret_cast<float4*>(p.z_out + row * (long long)b + g * 32 + h * 16);
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    zo[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                        __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
              }
              words[g] = w;
            }
          }
          if (row < p.n) {
            uint32_t* out = p.codes + row * p.W;
            for (int i = 0; i < wpad; ++i) out[i] = 0u;
            if (nb == 8 && wpad == 0) {
              reinterpret_cast<uint4*>(out)[0] = make_uint4(words[0], words[1], words[2], words[3]);
              reinterpret_cast<uint4*>(out)[1] = make_uint4(words[4], words[5], words[6], words[7]);
            } else {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                if (g < nb) out[wpad + g] = words[g];
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
        acc_phase ^= 1;
      }
    } else {
      // ---- flat L2 filter: the accumulator is x.q - |x|^2/2 - (|q|^2 - tq)/2, so a pair passes its
      // query's threshold (d2 <= tq) iff the accumulator is >= 0: a sign test like the hash epilogue.
      // Survivors (rare after the first chunk) are appended to the query's candidate buffer.
      for (long long t = 0; t < my_tiles; ++t) {
        const long long tile = blockIdx.x + t * gridDim.x;
        const long long jb = tile / row_tiles, rt = tile - jb * row_tiles;
        mbar_wait(acc_full, acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const long long row = rt * TILE_M + sub * UMMA_M + warp * 32 + lane;
          const bool rvalid = row < p.n;
          const uint32_t tbase = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(sub * b);
          const unsigned row_id = p.row_base + (unsigned)row;
          const long long q0 = jb * 256;
          uint32_t va[16], vb[16];                                // ping-pong: no register copies
          tmem_ld16(tbase, va);
#pragma unroll 1
          for (int g16 = 0; g16 < 2 * nb; g16 += 2) {
            tmem_ld16_nowait(tbase + (uint32_t)((g16 + 1) * 16), vb);
            unsigned m = rvalid ? nonneg_mask16(va) : 0u;
            if (m) SB_L2_APPEND(m, va, q0 + g16 * 16, row_id, p.tq, p.cand_buf, p.cand_cnt, p.cap);
            tmem_ld_wait();
            if (g16 + 2 < 2 * nb) tmem_ld16_nowait(tbase + (uint32_t)((g16 + 2) * 16), va);
            m = rvalid ? nonneg_mask16(vb) : 0u;
            if (m) SB_L2_APPEND(m, vb, q0 + (g16 + 1) * 16, row_id, p.tq, p.cand_buf, p.cand_cnt, p.cap);
            tmem_ld_wait();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
        acc_phase ^= 1;
      }
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

int pick_stages(int b, size_t* smem_bytes) {
  const size_t stage = 2 * (size_t)A_PART + 2 * (size_t)KCHUNKS * b * 16;
  const size_t tail = (2 * MAX_STAGES + 2) * 8 + 16;
  int s = (int)((220 * 1024 - tail) / stage);
  if (s > MAX_STAGES) s = MAX_STAGES;
  if (smem_bytes) *smem_bytes = (size_t)s * stage + tail;
  return s;
}

}  // namespace

namespace sb {

// Flat-L2 filter pass (flat_l2.cu): rows [0, n) of X against `col_blocks` blocks of 256
// query columns whose hi/lo images (same layout as the rotation image, b = 256, plus one
// synthetic K chunk carrying the thresholds) start at q_image + jb * (D + 16) * 512 words.
// Pairs with |x|^2 + |q|^2 - 2 x.q <= tq[query] are appended to cand_buf[query][cap]
// (key = d2 bits << 32 | row_base + row), counted in cand_cnt.
int tc_l2_filter(const float* X, int64_t n, int32_t D, int64_t ldx, const uint32_t* q_image, int col_blocks,
                 const float* xn, const float* tq, unsigned long long* cand_buf, int* cand_cnt, int cap,
                 unsigned row_base, int passes, cudaStream_t st) {
  TcParams p;
  p.X = X; p.n = n; p.D = D; p.ldx = ldx; p.mean = nullptr; p.row_div = nullptr;
  p.r_image = q_image; p.b = 256; p.codes = nullptr; p.W = 8; p.z_out = nullptr;
  size_t smem_bytes = 0;
  p.stages = pick_stages(256, &smem_bytes);
  p.tmem_cols = 512;
  p.col_blocks = col_blocks;
  p.image_block_words = (long long)(D + KC) * 256 * 2;         // D/16 data chunks + the synthetic chunk
  p.passes = passes;
  p.xn = xn; p.tq = tq; p.cand_buf = cand_buf; p.cand_cnt = cand_cnt; p.cap = cap; p.row_base = row_base;
  const long long tiles = ((n + TILE_M - 1) / TILE_M) * col_blocks;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  SB_CUDA_TRY(cudaFuncSetAttribute(itq_hash_tc_kernel<false, EPI_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem_bytes));
  ProfScope prof("l2_filter_tc_kernel", st);
  itq_hash_tc_kernel<false, EPI_L2><<<grid, THREADS, smem_bytes, st>>>(p);
  count_launch();
  return check_launch("l2_filter_tc_kernel");
}

int itq_hash_tc_supported(int64_t n, int32_t D, int64_t ldx, int32_t b, const float* X) {
  return n >= 1 && D >= KC && D % KC == 0 && b >= 32 && b <= 256 && b % 32 == 0 && ldx % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(X) & 15u) == 0;
}

}  // namespace sb

extern "C" {

size_t sb_itq_rotation_image_bytes(int32_t D, int32_t b) {
  if (D < KC || D % KC != 0 || b < 32 || b > 256 || b % 32 != 0) return 0;
  return (size_t)D * b * 4 * 2;
}

int sb_itq_rotation_image(const float* R, int32_t D, int32_t b, void* image_out, void* stream) {
  SB_REQUIRE(R != nullptr && image_out != nullptr, "sb_itq_rotation_image: NULL pointer");
  SB_REQUIRE(sb_itq_rotation_image_bytes(D, b) != 0,
             "sb_itq_rotation_image: tensor-core hashing needs D %% 16 == 0, 32 <= b <= 256, b %% 32 == 0 (D=%d b=%d)", D, b);
  SB_REQUIRE((reinterpret_cast<uintptr_t>(image_out) & 15u) == 0, "sb_itq_rotation_image: image must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)D * b;
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  sb::ProfScope prof("rotation_image_kernel", st);
  rotation_image_kernel<<<blocks, 256, 0, st>>>(R, D, b, reinterpret_cast<uint32_t*>(image_out));
  sb::count_launch();
  return sb::check_launch("rotation_image_kernel");
}

int sb_itq_hash_tc(const float* X, int64_t n, int32_t D, int64_t ldx, const float* mean, const void* r_image, int32_t b,
                   const float* row_div, uint32_t* codes_out, int32_t W, float* z_out, void* stream) {
  SB_REQUIRE(n >= 0 && D >= 1 && b >= 1, "sb_itq_hash_tc: need n>=0, D>=1, b>=1");
  SB_REQUIRE(ldx >= D, "sb_itq_hash_tc: ldx=%lld < D=%d", (long long)ldx, D);
  SB_REQUIRE(W * 32 >= b, "sb_itq_hash_tc: %d words cannot hold %d bits", W, b);
  SB_REQUIRE(r_image != nullptr && codes_out != nullptr, "sb_itq_hash_tc: NULL pointer");
  if (n == 0) return SB_OK;
  SB_REQUIRE(X != nullptr, "sb_itq_hash_tc: X is NULL");
  if (!sb::itq_hash_tc_supported(n, D, ldx, b, X) || (reinterpret_cast<uintptr_t>(r_image) & 15u) ||
      (mean && (reinterpret_cast<uintptr_t>(mean) & 15u)) || (reinterpret_cast<uintptr_t>(codes_out) & 15u) ||
      (z_out && (reinterpret_cast<uintptr_t>(z_out) & 15u))) {
    sb::set_error("sb_itq_hash_tc: needs D%%16==0, b%%32==0, 32<=b<=256, ldx%%4==0 and 16-byte aligned X, mean, image, outputs");
    return SB_ERR_UNSUPPORTED;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TcParams p;
  p.X = X; p.n = n; p.D = D; p.ldx = ldx; p.mean = mean; p.row_div = row_div;
  p.r_image = reinterpret_cast<const uint32_t*>(r_image);
  p.b = b; p.codes = codes_out; p.W = W; p.z_out = z_out;
  p.col_blocks = 1; p.image_block_words = 0;
  p.xn = nullptr; p.tq = nullptr; p.cand_buf = nullptr; p.cand_cnt = nullptr; p.cap = 0; p.row_base = 0;
  p.passes = 3;
  size_t smem_bytes = 0;
  p.stages = pick_stages(b, &smem_bytes);
  int cols = 32;
  while (cols < 2 * b) cols <<= 1;
  p.tmem_cols = cols;
  const long long tiles = (n + TILE_M - 1) / TILE_M;
  const int grid = (int)(tiles < sb::sm_count() ? tiles : sb::sm_count());
  sb::ProfScope prof("itq_hash_tc_kernel", st);
  if (row_div) {
    SB_CUDA_TRY(cudaFuncSetAttribute(itq_hash_tc_kernel<true, EPI_HASH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes));
    itq_hash_tc_kernel<true, EPI_HASH><<<grid, THREADS, smem_bytes, st>>>(p);
  } else {
    SB_CUDA_TRY(cudaFuncSetAttribute(itq_hash_tc_kernel<false, EPI_HASH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes));
    itq_hash_tc_kernel<false, EPI_HASH><<<grid, THREADS, smem_bytes, st>>>(p);
  }
  sb::count_launch();
  return sb::check_launch("itq_hash_tc_kernel");
}

}  // extern "C"
