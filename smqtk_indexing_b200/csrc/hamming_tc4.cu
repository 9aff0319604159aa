// Batched Hamming scan / top-k on the tensor cores, PACKED FP4 operands (reference: LinearHashIndex._nn,
// smqtk_indexing/impls/hash_index/linear.py:232-240 + hamming_distance utils/metrics.py:155).
//
// Same algorithm, data flow and bookkeeping as hamming_tc.cu (read its header first); what changes is the
// operand format: bits are mapped to +-1 in E2M1 (4 bits per element: +1.0 = 0x2, -1.0 = 0xA) and multiplied
// with tcgen05.mma kind::mxf4.block_scale -- K = 64 elements per MMA at the cycle count of kind::f8f6f4's K = 32
// (measured, tools/experiments/mxf4_probe2.cu: 137.6 cycles for 128 x 256 x 64 vs 137.6 for 128 x 256 x 32), i.e.
// HALF the tensor-pipe time for the same dot products, still exact: every product is +-1 and the FP32
// accumulator holds the integers.  A 256-bit code is ONE 128-byte SWIZZLE_128B row (4 MMAs, start address
// advancing 32 bytes), so a tile has one K group instead of two.
//   * Block scaling cannot be switched off, so it is put to use (below).  The scale factors live in tensor memory;
//     every SF window is filled with ONE value in every byte of every lane with tcgen05.st once per CTA, so the
//     (sub-partition replicated) SF layout never matters.  Instruction descriptor: a/b format = 1
//     (MXF4Format::E2M1; 5 is the kind::mxf8f6f4 code and raises "illegal instruction"), scale format UE8M0.
//   * Accumulators are FP32 only, and every accumulator leaves tensor memory through tcgen05.ld at 64 B/clk/SM:
//     with one query per column that read sets the pace (measured: profiles/r2_ham_fp4_full.md).  So every
//     accumulator column carries THREE queries in 8-bit fields (all 24 bits of an FP32 integer):
//         X = u_a + 256 w_b + 65536 u_c,      acc = X - 2^23,
//         u_a = tq_a - d_a + beta,  w_b = d_b - tq_b + gamma,  u_c = tq_c - d_c + beta
//     with T = the largest threshold of the batch when the chunk starts (collected by the compaction kernel behind
//     the previous chunk), beta = (254 - T) & ~1 and gamma = T + 1: "d <= tq" is u >= beta / w <= gamma, the
//     same constants for every column.  Query operands are +-0.5 (a dot product is K/2 - d; the b query has its
//     signs inverted), the b and c MMAs use scale-factor windows holding 2^8 and 2^16 (uniform per MMA, so the SF
//     layout still does not matter), and the thresholds ride in one extra K step per field.  4/3 bytes per pair.
//     Decoding: v = bits(acc +rz 1.5 * 2^24) = 0x4B800000 | (X >> 1); VIMNMX3.U16x2 keeps a running min of v << 1
//     (low half [w_b : u_a >> 1], high half [exponent : u_c]) and a running max of v << 9 (low half [u_a >> 1],
//     high half [u_c : w_b]) over a row's 32 columns: 4 instructions per accumulator, three compares per group.
//     A field can leave its window only for a distance far above the threshold (d > 254 - (T - tq)); the fields
//     are oriented so that this can only ADD survivors (the exact re-check drops them): u_a < 0 borrows from w_b
//     (smaller = more likely to pass), w_b > 255 carries into u_c (larger = more likely to pass), and u_c < 0 drops
//     the exponent of v, which the high half of the running min sees -- the row is then re-checked for the whole
//     group.  T > 252 (no room for the windows) raises the overflow flag: the caller re-runs the XOR/POPC scan.
//   * TMEM: accumulator 0 = columns [0, C), C <= 192; scale factors in [224, 256): 2^0 (SFA at 224, SFB at 232),
//     2^8 at 240, 2^16 at 248; accumulator 1 = [256, 256 + C).  A block = C columns = 3 C queries: query
//     jb * 3C + h * C + n sits in column n as field h.
//   * The seed chunk (every pair kept) is a plain XOR/POPC kernel: it is a few hundred rows.
//   * Bit -> nibble expansion is two instructions per 8 elements (rows: +-1.0; queries: ^ 0x33333333 -> +-0.5):
//     out = ((w << s) & 0x88888888) | 0x22222222
//     for s = 3, 2, 1, 0 takes every 4th bit of a code word as the sign of 8 consecutive elements.  This permutes
//     the elements inside a row, identically for table rows and query rows -- a dot product does not care.
//   * Threshold steps: A_syn = 64 x (+1), B_syn = up to 64 E2M1 slots (values 6, 4, 3, 2, 1) summing to the field's offset.
//   * Survivor groups are 32 columns = 32 queries of each field, flagged separately; every 32-query group has its own
//     re-check list of rows (batches of <= 1024 queries with k < 32: one list of (row, group) entries, see make_plan).
// Warp roles: 0-15 epilogue (lane quadrant x tile parity x column half), 16 MMA issue + TMEM alloc, 17 B loader,
// 18-21 producers (one warp per stage of the A ring).
#include <cuda_fp16.h>

#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

using namespace tcptx;

namespace {

constexpr int TM = 128;                     // rows per tile (UMMA M)
constexpr int QB = 192;                     // most columns per block (UMMA N): two block images of 480 B per column in shared memory
constexpr int QPC = 3;                      // queries per accumulator column
constexpr int A_GROUP = TM * 128;           // 16 KB: 128 rows x 256 E2M1 (8 code words)
// one query block in shared memory: qb rows x 128 B (SWIZZLE_128B) + B_syn 2 x qb x 16 B (no swizzle), rounded up
// to the 1024-byte swizzle atom -- 38 KB at qb = 240; two of them (the next block streams in during the current pass)
constexpr int A_SYN = 2 * TM * 16;          // 4 KB
constexpr int SF_COL = 224;                 // TMEM columns [224, 240): scale factors 2^0 (SFA at +0, SFB at +8)
constexpr int SF8_COL = 240;                // [240, 248): 2^8 (SFB of the b-field MMAs)
constexpr int SF16_COL = 248;               // [248, 256): 2^16 (SFB of the c-field MMAs)
constexpr float DECODE_MAGIC = 25165824.0f; // 1.5 * 2^24: bits(acc +rz this) = 0x4B800000 | ((acc + 2^23) >> 1)
constexpr int T_MAX = 252;                  // largest threshold the 8-bit windows take
constexpr int ACC1_COL = 256;               // second accumulator
constexpr int MAX_STAGES = 4;               // A ring: 16 KB per stage; 4 stages when the two block images leave room, else 2 (HamTc4Params::stages)
constexpr int GRAN = 32;                    // rows per granule of the visiting order
// 16 epilogue warps: (TMEM lane quadrant) x (accumulator buffer = tile parity) x (column half).  tcgen05.wait::ld
// waits for ALL of a thread's loads, so tensor-memory latency can only be hidden by other warps -- with FP32
// accumulators (twice the registers per column of the FP8 kernel's packed halves) and half the MMA time per
// tile, four epilogue warps were the bottleneck (measured: 6.8 ms vs the FP8 kernel's 6.3 ms).
constexpr int EPI_WARPS = 16, EPI_PER_BUF = 8, MMA_WARP = 16, B_WARP = 17, PROD_WARP0 = 18, PROD_WARPS = MAX_STAGES;
constexpr int THREADS = (PROD_WARP0 + PROD_WARPS) * 32;   // 704
constexpr int GROWTH = 4;                   // rows of a chunk = 3 x the rows before it: ~3 (k + ties) survivors per query
constexpr int GROWTH_SMALL_Q = 8;           // few queries: every chunk's launch / fill / drain weighs more than the
constexpr int SMALL_Q = 1024;               // extra survivors -- 6 chunks of 7 (k + ties) instead of 9 of 3 (k + ties)
constexpr int CP_THREADS = 256;            // 7-8 compaction CTAs per SM: most queries hold a few hundred keys

struct HamTc4Params {
  const uint32_t* db;          // u32[U][W]
  long long U;
  int W;                       // code words per row
  long long vg0, vg1;          // virtual granules of this chunk
  long long NG, P;             // physical granule = (virtual * P) mod NG
  int col_blocks, cb_per;      // query blocks in total / per blockIdx.y
  const unsigned char* image;  // per block: G x B_GROUP (SW128) then B_SYN
  const int* tq;               // thresholds (Hamming distance) per query slot
  const int* tqmax;            // T: upper bound of the thresholds (device scalar)
  unsigned long long* recheck; // per 32-query group: recheck_cap rows to re-check
  int qb, b_block;             // columns per block (multiple of 32, <= 192; 3 * qb queries), bytes of one block image
  int* recheck_cnt;            // entries per 32-query group (by_group) / of the one list
  int recheck_cap;
  int by_group;                // per-group lists (many queries) or one list of (row << 24 | 31 << 19 | group first query / 16) entries
  unsigned long long* cand_buf;
  int* cand_cnt;
  int cap;
  long long idx_base;
  int stages;                  // stages of the A ring in use (2 or 4)
};

// 32 code bits -> 32 E2M1 elements (16 bytes): +1.0 = 0x2, -1.0 = 0xA.  Output word s holds the bits
// s, s + 4, s + 8, ... of the code word as the signs of 8 consecutive elements (an element permutation inside
// the row, the same for every row of both operands).
__device__ __forceinline__ uint4 expand_word4(uint32_t w) {
  uint4 o;
  o.x = ((w << 3) & 0x88888888u) | 0x22222222u;
  o.y = ((w << 2) & 0x88888888u) | 0x22222222u;
  o.z = ((w << 1) & 0x88888888u) | 0x22222222u;
  o.w = (w & 0x88888888u) | 0x22222222u;
  return o;
}
// byte offset of 16-byte chunk c of row r inside a [rows x 128 B] SWIZZLE_128B K-major group
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
}

// D[tmem] (+)= A[smem] . B[smem]^T, kind::mxf4 block-scaled (E2M1 x E2M1 -> F32, K = 64), issued by one thread.
__device__ __forceinline__ void umma_fp4(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                         uint32_t tsfa, uint32_t tsfb) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tsfa), "r"(tsfb)
      : "memory");
}

// One re-check entry = (row, group of up to 32 queries starting at `qg0`): exact distances from the packed codes
// (see hamming_tc.cu for why survivors are not decoded from the accumulator registers).
template <int W>
__device__ __forceinline__ void ham4_recheck_group(const uint32_t* __restrict__ db, const uint32_t* __restrict__ qcodes, int Q,
                                                   long long row, unsigned long long row_key, int qg0, int width,
                                                   const int* __restrict__ tq, unsigned long long* cand_buf, int* cand_cnt,
                                                   int cap, int lane) {
  uint32_t x[W];
#pragma unroll
  for (int w = 0; w < W; ++w) x[w] = __ldg(db + row * W + w);
  const int qg = qg0 + lane;
  if (lane < width && qg < Q) {
    int d = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) d += __popc(x[w] ^ __ldg(qcodes + (long long)qg * W + w));
    if (d <= tq[qg]) {
      const int slot = atomicAdd(cand_cnt + qg, 1);
      if (slot < cap) cand_buf[(long long)qg * cap + slot] = ((unsigned long long)(unsigned)d << 40) | row_key;
    }
  }
}

// W = code words per row (1, 2, 4, 8); 3 x (W / 2 MMAs of K = 64 elements + the threshold step) per tile.
// ONE query block is resident at a time, its successor streams into the second buffer.  The expansion of a table
// tile is two instructions per 8 elements, so re-expanding it for every query block is cheap and the table stays
// L2 / HBM resident.
template <int W, bool BY_GROUP>
__global__ void __launch_bounds__(THREADS, 1) ham_filter_fp4_kernel(const HamTc4Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms: align by hand
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int KSTEPS = (W >= 2) ? W / 2 : 1;                    // 64 elements = 2 code words per MMA
  constexpr uint32_t a_stage = (uint32_t)A_GROUP;
  const int qb = p.qb;                                           // columns per block (multiple of 32, <= 192): 3 * qb queries
  const uint32_t b_block = (uint32_t)p.b_block;                  // bytes of one block image (data rows + B_syn), 1024-aligned
  // layout (offsets are multiples of 1024): [B buffer 0][B buffer 1][A_syn][A ring][barriers, thresholds]
  unsigned char* s_b = smem;
  unsigned char* s_asyn = s_b + 2 * b_block;
  unsigned char* s_a = s_asyn + A_SYN;
  const int stages = p.stages;                                   // 2 or 4
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_a + (size_t)stages * a_stage);
  const uint32_t a_full = smem_u32(bars), a_empty = a_full + MAX_STAGES * 8;
  const uint32_t acc_full = a_empty + MAX_STAGES * 8, acc_empty = acc_full + 16;
  const uint32_t b_full = acc_empty + 16, b_empty = b_full + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 8);

  const long long n_gran = p.vg1 - p.vg0;
  const long long n_tiles = (n_gran + 3) / 4;
  const long long my_tiles = (n_tiles > (long long)blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int jb0 = blockIdx.y * p.cb_per;
  const int jb1 = min(p.col_blocks, jb0 + p.cb_per);

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(a_full + s * 8, 1); mbar_init(a_empty + s * 8, 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full + b * 8, 1);
      mbar_init(acc_empty + b * 8, EPI_PER_BUF);
      mbar_init(b_full + b * 8, 1);
      mbar_init(b_empty + b * 8, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // A_syn = +1.0 in every E2M1 element; codes narrower than 256 bits leave K chunks of the A ring untouched: zero
  // the ring once (a zero nibble is +0.0)
  for (int i = threadIdx.x; i < A_SYN / 4; i += THREADS) reinterpret_cast<uint32_t*>(s_asyn)[i] = 0x22222222u;
  if (W < 8)
    for (int i = threadIdx.x; i < (int)(stages * a_stage / 16); i += THREADS)
      reinterpret_cast<uint4*>(s_a)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // UE8M0 block scales, uniform per window (whatever layout the MMA reads): 2^0 = 0x7F in every byte of every lane
  // of columns [224, 240), 2^8 = 0x87 in [240, 248), 2^16 = 0x8F in [248, 256)
  if (warp < 4) {
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(SF_COL + c);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(c < 16 ? 0x7F7F7F7Fu : (c < 24 ? 0x87878787u : 0x8F8F8F8Fu))
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == B_WARP) {
    // =========================== B: the next query block streams in while the current one is used ===========
    if (lane == 0) {
      int it = 0;
      for (int jb = jb0; jb < jb1; ++jb, ++it) {
        const int sb = it & 1;
        mbar_wait(b_empty + sb * 8, ((it >> 1) & 1) ^ 1);
        const unsigned char* src = p.image + (size_t)jb * b_block;
        mbar_expect_tx(b_full + sb * 8, b_block);
        for (uint32_t o = 0; o < b_block; o += 1024) tma_bulk_g2s(smem_u32(s_b + sb * b_block + o), src + o, 1024, b_full + sb * 8);
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issue ===========================
    if (lane == 0) {
      // block-scaled descriptor: A = B = E2M1 (MXF4Format 1) at bits 7 / 10, both K-major, N at 17, scale
      // format UE8M0 (bit 23), M at 24, scale-factor ids 0, K = 64 (bit 31 = 0); accumulator FP32
      const uint32_t idesc = (1u << 7) | (1u << 10) | ((uint32_t)(qb >> 3) << 17) | (1u << 23) | ((uint32_t)(TM >> 4) << 24);
      const uint32_t tsfa = tmem_base + (uint32_t)SF_COL, tsfb = tmem_base + (uint32_t)(SF_COL + 8);
      const uint32_t tsfb8 = tmem_base + (uint32_t)SF8_COL;      // the b field's products are scaled by 2^8,
      const uint32_t tsfb16 = tmem_base + (uint32_t)SF16_COL;    // the c field's by 2^16
      const uint64_t asyn_desc = umma_desc(smem_u32(s_asyn), TM * 16, 128);
      const uint64_t a_desc0 = umma_desc_sw128(smem_u32(s_a));
      int stage = 0, it = 0;
      uint32_t phase = 0;
      long long t = 0;                                           // tiles issued (accumulator buffer = t & 1)
      for (int jb = jb0; jb < jb1; ++jb, ++it) {
        const int sb = it & 1;
        mbar_wait(b_full + sb * 8, (it >> 1) & 1);
        tc_fence_after();
        // block image: 3 x [qb rows x 128 B] (fields a, b, c), then 3 x B_syn [2 x qb x 16 B]
        const uint32_t bb = smem_u32(s_b) + sb * b_block;
        uint64_t b_desc[QPC], bsyn_desc[QPC];
#pragma unroll
        for (int h = 0; h < QPC; ++h) {
          b_desc[h] = umma_desc_sw128(bb + (uint32_t)(h * qb) * 128u);
          bsyn_desc[h] = umma_desc(bb + (uint32_t)qb * (128u * QPC) + (uint32_t)(h * qb) * 32u, (uint32_t)qb * 16u, 128);
        }
        for (long long i = 0; i < my_tiles; ++i, ++t) {
          const int buf = (int)(t & 1);
          mbar_wait(a_full + stage * 8, phase);
          mbar_wait(acc_empty + buf * 8, (uint32_t)((t >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + (uint64_t)(((uint32_t)stage * a_stage) >> 4);
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * ACC1_COL);
#pragma unroll
          for (int h = 0; h < QPC; ++h) {
            const uint32_t sfb = h == 0 ? tsfb : (h == 1 ? tsfb8 : tsfb16);
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks)
              umma_fp4(d_tmem, a_desc + (uint64_t)((ks * 32) >> 4), b_desc[h] + (uint64_t)((ks * 32) >> 4), idesc, (h | ks) ? 1u : 0u,
                       tsfa, sfb);
            umma_fp4(d_tmem, asyn_desc, bsyn_desc[h], idesc, 1u, tsfa, sfb);
          }
          umma_commit(acc_full + buf * 8);
          umma_commit(a_empty + stage * 8);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(b_empty + sb * 8);                          // this B buffer may be replaced once these MMAs retire
      }
    }
  } else if (warp >= PROD_WARP0) {
    // =========================== A: packed codes -> +-1 E2M1, swizzled ===========================
    // One warp per ring stage: warp pw expands the WHOLE tiles t = pw (mod MAX_STAGES) -- lane = row of each of the
    // tile's four 32-row granules, all W words -- so that the generic -> async proxy fence (several hundred cycles:
    // it was more than half of a producer's time when every producer thread touched every tile, and with ~1000-cycle
    // tiles it set the pace) is paid once per warp and MAX_STAGES tiles, not once per tile.
    const int pw = warp - PROD_WARP0;
    const unsigned long long NGu = (unsigned long long)p.NG;
    const long long Pm = (long long)((unsigned long long)p.P % NGu);
    uint32_t cw[4][W];                                           // the warp's next tile: granule g, word j of this lane's row
    unsigned char* dst = s_a + (size_t)pw * a_stage;
    long long t0 = 0;                                            // global tile counter at the start of this block's pass
    auto load = [&](long long i) {                               // tile i of the pass: granules (blockIdx.x + i * gridDim.x) * 4 + g
      const long long vg = p.vg0 + ((long long)blockIdx.x + i * gridDim.x) * 4;
      long long pg = (long long)(((unsigned long long)vg * (unsigned long long)p.P) % NGu);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const long long row = pg * GRAN + lane;
        const bool valid = (vg + g < p.vg1) && row < p.U;
        const uint32_t* src = p.db + row * W;
        if (W == 8) {
          uint4 x0 = make_uint4(0u, 0u, 0u, 0u), x1 = x0;
          if (valid) { x0 = __ldg(reinterpret_cast<const uint4*>(src)); x1 = __ldg(reinterpret_cast<const uint4*>(src) + 1); }
          cw[g][0] = x0.x; cw[g][1 % W] = x0.y; cw[g][2 % W] = x0.z; cw[g][3 % W] = x0.w;
          cw[g][4 % W] = x1.x; cw[g][5 % W] = x1.y; cw[g][6 % W] = x1.z; cw[g][7 % W] = x1.w;
        } else if (W == 4) {
          uint4 x0 = make_uint4(0u, 0u, 0u, 0u);
          if (valid) x0 = __ldg(reinterpret_cast<const uint4*>(src));
          cw[g][0] = x0.x; cw[g][1 % W] = x0.y; cw[g][2 % W] = x0.z; cw[g][3 % W] = x0.w;
        } else if (W == 2) {
          uint2 x0 = make_uint2(0u, 0u);
          if (valid) x0 = __ldg(reinterpret_cast<const uint2*>(src));
          cw[g][0] = x0.x; cw[g][1 % W] = x0.y;
        } else {
          cw[g][0] = valid ? __ldg(src) : 0u;
        }
        pg += Pm;
        if (pg >= p.NG) pg -= p.NG;
      }
    };
    for (int jb = jb0; jb < jb1 && pw < stages; ++jb, t0 += my_tiles) {
      const long long i_first = (pw - t0) & (stages - 1);        // first tile of this pass that lands in this warp's stage
      uint32_t phase = (uint32_t)(((t0 + i_first) / stages) & 1);
      if (i_first < my_tiles) load(i_first);
      for (long long i = i_first; i < my_tiles; i += stages) {
        mbar_wait(a_empty + pw * 8, phase ^ 1);
        phase ^= 1u;
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int j = 0; j < W; ++j)                              // (rows past the table expand to some code: the epilogue masks them)
            *reinterpret_cast<uint4*>(dst + sw128_off(g * 32 + lane, j)) = expand_word4(cw[g][j]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full + pw * 8);
        if (i + stages < my_tiles) load(i + stages);               // in flight while the ring turns
      }
    }
  } else if (warp < EPI_WARPS) {
    // =========================== epilogue: window tests, survivors queued ===========================
    const int ew = warp & 3;                                     // TMEM lane quadrant = granule of the tile
    const int buf = (warp >> 2) & 1;                             // accumulator buffer: this warp takes the tiles t = buf (mod 2)
    constexpr int GRP = 3;                                       // most 32-column survivor groups per warp (qb <= 192)
    const int split = ((qb + 63) >> 6) << 5;                     // columns of the first half (whole groups)
    const int col0 = (warp >> 3) * split;                        // this warp's columns: [0, split) or [split, qb)
    const int ngrp = ((warp >> 3) ? (qb - split) : split) >> 5;
    const long long vg_base = p.vg0 + (long long)blockIdx.x * 4 + ew;
    const long long vg_step2 = 8ll * gridDim.x;                  // this warp sees every second tile
    const unsigned long long NGu = (unsigned long long)p.NG;
    const long long pg_step2 = (long long)(((unsigned long long)vg_step2 * (unsigned long long)p.P) % NGu);
    const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * ACC1_COL + col0);
    // window constants of this batch (see the header): d <= tq  <=>  u_a >= beta, w_b <= gamma, u_c >= beta
    const int T = min(__ldg(p.tqmax), T_MAX);
    const uint32_t beta = (uint32_t)((254 - T) & ~1), gamma = (uint32_t)(T + 1);
    const uint32_t a_min = (beta >> 1) << 9;                     // low half of max(v << 9) = (u_a >> 1) << 9
    const uint32_t c_min = beta << 8;                            // high half of max(v << 9) = [u_c : w_b]
    const uint32_t b_lim = (gamma + 1u) << 8;                    // low half of min(v << 1) = [w_b : u_a >> 1 : 0]
    long long t0 = 0;                                            // global tile counter at the start of this block's pass
    for (int jb = jb0; jb < jb1; ++jb, t0 += my_tiles) {
      const int q0 = jb * QPC * qb + col0;                       // first a-field query of this warp's columns; field h: + h * qb
      const long long i_first = ((t0 & 1) == buf) ? 0 : 1;       // first tile of this pass that lands in this warp's buffer
      long long vg = vg_base + i_first * 4 * (long long)gridDim.x;
      long long pg = (long long)(((unsigned long long)vg * (unsigned long long)p.P) % NGu);
      uint32_t par = (uint32_t)(((t0 + i_first) >> 1) & 1);
      for (long long i = i_first; i < my_tiles; i += 2) {
        const long long row = pg * GRAN + lane;
        const bool rvalid = (vg < p.vg1) && (row < p.U);
        mbar_wait(acc_full + buf * 8, par);
        par ^= 1u;
        tc_fence_after();
        unsigned hit[GRP][QPC];                                    // lanes with a surviving query of field h in each 32-column group
#pragma unroll
        for (int g = 0; g < GRP; ++g) {
#pragma unroll
          for (int h = 0; h < QPC; ++h) hit[g][h] = 0u;
          if (g < ngrp) {                                          // warp-uniform
            uint32_t va[32];
            tmem_ld32_nowait(tbase + (uint32_t)(32 * g), va);
            tmem_ld_wait();
            uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const uint32_t v0 = __float_as_uint(__fadd_rz(__uint_as_float(va[j]), DECODE_MAGIC));
              const uint32_t v1 = __float_as_uint(__fadd_rz(__uint_as_float(va[j + 1]), DECODE_MAGIC));
              mn = __vimin3_u16x2(mn, v0 << 1, v1 << 1);
              mx = __vimax3_u16x2(mx, v0 << 9, v1 << 9);
            }
            const bool under = (mn >> 16) < 0x9700u;               // a c field below its window: the other fields are unreadable
            hit[g][0] = __ballot_sync(0xffffffffu, rvalid && (under || (mx & 0xffffu) >= a_min));
            hit[g][1] = __ballot_sync(0xffffffffu, rvalid && (under || (mn & 0xffffu) < b_lim));
            hit[g][2] = __ballot_sync(0xffffffffu, rvalid && (under || (mx >> 16) >= c_min));
          }
        }
        // the accumulator buffer is free: let the MMAs of tile t + 2 start, THEN pay for the appends
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + buf * 8);
        unsigned any = 0u;
#pragma unroll
        for (int g = 0; g < GRP; ++g)
#pragma unroll
          for (int h = 0; h < QPC; ++h) any |= hit[g][h];
        if (any && !BY_GROUP) {                                    // one list, one atomic per (warp, tile) with survivors
          int total = 0;
#pragma unroll
          for (int g = 0; g < GRP; ++g)
#pragma unroll
            for (int h = 0; h < QPC; ++h) total += __popc(hit[g][h]);
          int base = 0;
          if (lane == 0) base = atomicAdd(p.recheck_cnt, total);
          base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
          for (int g = 0; g < GRP; ++g) {
#pragma unroll
            for (int h = 0; h < QPC; ++h) {
              const unsigned m = hit[g][h];
              if ((m >> lane) & 1u) {
                const int slot = base + __popc(m & ((1u << lane) - 1u));
                if (slot < p.recheck_cap)
                  p.recheck[slot] = ((unsigned long long)row << 24) | (31ull << 19) | (unsigned long long)((q0 + h * qb + 32 * g) >> 4);
              }
              base += __popc(m);
            }
          }
        } else if (any) {                                          // rare: survivors in this (warp, tile)
          // one slot reservation per (group, field) WITH survivors, all issued at once: lane c reserves for combination
          // c = g * 3 + h (serial atomics would cost a global round trip each)
          int mine_cnt = 0;
#pragma unroll
          for (int g = 0; g < GRP; ++g)
#pragma unroll
            for (int h = 0; h < QPC; ++h) mine_cnt = (lane == g * QPC + h) ? __popc(hit[g][h]) : mine_cnt;
          int mine_base = 0;
          if (lane < GRP * QPC && mine_cnt)
            mine_base = atomicAdd(p.recheck_cnt + ((q0 + (lane % QPC) * qb + 32 * (lane / QPC)) >> 5), mine_cnt);
#pragma unroll
          for (int g = 0; g < GRP; ++g) {
#pragma unroll
            for (int h = 0; h < QPC; ++h) {
              const unsigned m = hit[g][h];
              const int base = __shfl_sync(0xffffffffu, mine_base, g * QPC + h);
              if ((m >> lane) & 1u) {
                const int slot = base + __popc(m & ((1u << lane) - 1u));
                if (slot < p.recheck_cap)
                  p.recheck[(size_t)((q0 + h * qb + 32 * g) >> 5) * p.recheck_cap + slot] = (unsigned long long)row;
              }
            }
          }
        }
        vg += vg_step2;
        pg += pg_step2;
        if (pg >= p.NG) pg -= p.NG;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Few queries (<= 1024: the per-rank share of a batch on 4-8 GPUs): ONE unordered list of (row, 32-query group)
// entries, one warp per entry -- every entry re-reads its group's query codes (1 KB from L2), but the kernel is a
// few microseconds shorter than the per-group one when there are only a handful of groups.
template <int W>
__global__ void __launch_bounds__(256)
ham4_recheck_flat_kernel(const uint32_t* __restrict__ db, const uint32_t* __restrict__ qcodes, int Q, long long idx_base,
                    const unsigned long long* __restrict__ list, const int* __restrict__ list_cnt, int list_cap,
                    const int* __restrict__ tq, unsigned long long* __restrict__ cand_buf, int* __restrict__ cand_cnt, int cap,
                    int* __restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const int raw = *list_cnt;
  if (raw > list_cap && blockIdx.x == 0 && threadIdx.x == 0) *overflow = 1;
  const int n = min(raw, list_cap);
  const int warps = gridDim.x * (blockDim.x >> 5);
  for (int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < n; e += warps) {
    const unsigned long long ent = list[e];
    const long long row = (long long)(ent >> 24);
    const int qg0 = (int)(ent & 0x7ffffull) * 16;
    const int width = (int)((ent >> 19) & 31ull) + 1;
    ham4_recheck_group<W>(db, qcodes, Q, row, (unsigned long long)(idx_base + row), qg0, width, tq, cand_buf, cand_cnt, cap, lane);
  }
}

// Re-check, one CTA column per 32-query group (blockIdx.x = group, blockIdx.y splits the group's list): every lane
// keeps ITS query of the group in registers for the whole kernel, so an entry costs one row read instead of the row
// plus the group's 32 query codes.  A warp takes 32 entries at a time: lane l loads entry l's row, then the 32 rows
// are broadcast one after the other with shuffles and every lane tests its own query.  (Same scheme as
// ham_recheck_kernel of hamming_tc.cu.)
template <int W>
__global__ void __launch_bounds__(256)
ham4_recheck_kernel(const uint32_t* __restrict__ db, const uint32_t* __restrict__ qcodes, int Q, long long idx_base,
                    const unsigned long long* __restrict__ list, const int* __restrict__ list_cnt, int list_cap,
                    const int* __restrict__ tq, unsigned long long* __restrict__ cand_buf, int* __restrict__ cand_cnt, int cap,
                    int* __restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const int grp = blockIdx.x;
  const int raw = list_cnt[grp];
  if (raw > list_cap && blockIdx.y == 0 && threadIdx.x == 0) *overflow = 1;
  const int n = min(raw, list_cap);
  if ((blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32 >= n) return;   // no entries for this warp: leave before any load
  const int qg = grp * 32 + lane;
  const bool qvalid = qg < Q;
  uint32_t qw[W];
#pragma unroll
  for (int w = 0; w < W; ++w) qw[w] = qvalid ? __ldg(qcodes + (long long)qg * W + w) : 0u;
  const int my_tq = qvalid ? tq[qg] : -1;
  const unsigned long long* mine = list + (size_t)grp * list_cap;
  const int warps = gridDim.y * (blockDim.x >> 5);
  for (int e0 = (blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; e0 < n; e0 += warps * 32) {
    const int cnt = min(32, n - e0);
    long long row = 0;
    uint32_t x[W];
#pragma unroll
    for (int w = 0; w < W; ++w) x[w] = 0u;
    if (lane < cnt) {
      row = (long long)mine[e0 + lane];
      const uint32_t* src = db + row * W;
      if (W >= 4) {
#pragma unroll
        for (int w = 0; w < W; w += 4) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + w));
          x[w] = v.x; x[(w + 1) % W] = v.y; x[(w + 2) % W] = v.z; x[(w + 3) % W] = v.w;
        }
      } else {
#pragma unroll
        for (int w = 0; w < W; ++w) x[w] = __ldg(src + w);
      }
    }
    for (int i = 0; i < cnt; ++i) {
      int d = 0;
#pragma unroll
      for (int w = 0; w < W; ++w) d += __popc(__shfl_sync(0xffffffffu, x[w], i) ^ qw[w]);
      const long long ri = __shfl_sync(0xffffffffu, row, i);
      if (d <= my_tq) {
        const int slot = atomicAdd(cand_cnt + qg, 1);
        if (slot < cap) cand_buf[(long long)qg * cap + slot] = ((unsigned long long)(unsigned)d << 40) | (unsigned long long)(idx_base + ri);
      }
    }
  }
}

// Query codes -> per-block image: 3 x [qb rows x 128 B] (fields a, b, c) in the SWIZZLE_128B K-major order (same
// bit -> nibble order as the table rows; values +-0.5, the b field with inverted signs), then the three B_syn areas;
// query slots past Q and K chunks past the code are zero (the image is cleared first).  One thread per (query, word).
__global__ void ham4_query_image_kernel(const uint32_t* __restrict__ q, int Q, int W, int qb, int b_block,
                                        unsigned char* __restrict__ img) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Q * W) return;
  const int qi = (int)(i / W), j = (int)(i % W);
  const int jb = qi / (QPC * qb), r = qi % (QPC * qb);
  const int h = r / qb, n = r % qb;
  uint4 o = expand_word4(q[(long long)qi * W + j]);
  const uint32_t fix = (h == 1) ? 0xbbbbbbbbu : 0x33333333u;    // 1.0 (0x2) -> 0.5 (0x1); field b: sign flipped as well
  o.x ^= fix; o.y ^= fix; o.z ^= fix; o.w ^= fix;
  *reinterpret_cast<uint4*>(img + (size_t)jb * b_block + (size_t)h * qb * 128 + sw128_off(n, j)) = o;
}

// B_syn of every query slot: 64 E2M1 slots that sum to the field's offset (|s| <= 381: at most 63 slots of 6 plus
// the remainder), see the header:  a: tq - K/2 + beta,  b: gamma - tq + K/2,  c: tq - K/2 + beta - 128.
// Padding slots (their data nibbles are zero) sit inside their window on the failing side.
// One thread per (slot, 8-element word): the kernel runs between every two chunks, its latency is on the critical path.
__global__ void ham4_threshold_image_kernel(int Q, int cols, int K, int qb, int b_block, const int* __restrict__ tq,
                                            const int* __restrict__ tqmax, int* __restrict__ tqmax_next,
                                            unsigned char* __restrict__ img, int* __restrict__ list_cnt,
                                            int* __restrict__ overflow) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int col = tid >> 3, wd = tid & 7;                       // query slot, word (8 slots of 4 bits) of its 64 B_syn slots
  const int T_raw = *tqmax;
  if (col < cols && wd == 0 && (col & 31) == 0) list_cnt[col >> 5] = 0;   // the re-check lists restart with every chunk
  if (tid == 0) {
    *tqmax_next = 0;                                             // the compaction behind this chunk collects the next T here
    if (T_raw > T_MAX) *overflow = 1;                            // no room for the windows: the caller falls back
  }
  if (col >= cols) return;
  const int T = min(T_raw, T_MAX);
  const int beta = (254 - T) & ~1, gamma = T + 1;
  const int jb = col / (QPC * qb), r = col % (QPC * qb);
  const int h = r / qb, n = r % qb;
  int s;
  if (col < Q) {
    const int t = min(tq[col], T);
    s = (h == 1) ? (gamma - t + K / 2) : (t - K / 2 + beta - (h == 2 ? 128 : 0));
  } else {
    s = (h == 0) ? 0 : (h == 1 ? 255 : -128);
  }
  const uint32_t sign = s < 0 ? 0x8u : 0u;
  const int mag = s < 0 ? -s : s;
  const int n6 = mag / 6, rem = mag - 6 * n6;
  // E2M1 magnitude codes: 1.0 = 2, 2.0 = 4, 3.0 = 5, 4.0 = 6, 6.0 = 7; remainder 5 = 4 + 1 (two slots)
  const uint32_t rem_code[6] = {0u, 2u, 4u, 5u, 6u, 6u};
  uint32_t v = 0u;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const int slot = wd * 8 + b;
    uint32_t nib = 0u;
    if (slot < n6) nib = 7u | sign;
    else if (slot == n6 && rem) nib = rem_code[rem] | sign;
    else if (slot == n6 + 1 && rem == 5) nib = 2u | sign;
    v |= nib << (4 * b);
  }
  // B_syn of a field: two K-chunk planes of [qb x 16 B]; word wd of the 64 slots = plane wd / 4, word wd % 4
  unsigned char* base = img + (size_t)jb * b_block + (size_t)qb * (128 * QPC) + (size_t)h * qb * 32;
  reinterpret_cast<uint32_t*>(base + (size_t)(wd >> 2) * qb * 16 + n * 16)[wd & 3] = v;
}

__global__ void ham4_init_kernel(int cols, int K, int* __restrict__ tq, int* __restrict__ cnt, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { flags[0] = 0; flags[2] = 0; flags[3] = 0; }      // overflow flag, T (largest threshold) of odd / even chunks
  if (i >= cols) return;
  tq[i] = K;
  cnt[i] = 0;
}

// Seed chunk: buf[query][virtual row] = key for EVERY pair of the first `n_vrows` rows of the visiting order
// (rows past the table: empty keys, they sort last).  Thread = one row, 32 queries per blockIdx.y.
template <int W>
__global__ void __launch_bounds__(128)
ham4_seed_kernel(const uint32_t* __restrict__ db, long long U, long long NG, long long P, int n_vrows, const uint32_t* __restrict__ q,
                 int Q, long long idx_base, unsigned long long* __restrict__ cand_buf, int* __restrict__ cand_cnt, int cap) {
  const int vrow = blockIdx.x * blockDim.x + threadIdx.x;
  const int q0 = blockIdx.y * 32;
  if (vrow == 0)
    for (int j = 0; j < 32 && q0 + j < Q; ++j) cand_cnt[q0 + j] = n_vrows;
  if (vrow >= n_vrows) return;
  const long long pg = (long long)(((unsigned long long)(vrow / GRAN) * (unsigned long long)P) % (unsigned long long)NG);
  const long long row = pg * GRAN + (vrow % GRAN);
  const bool valid = row < U;
  uint32_t x[W];
#pragma unroll
  for (int w = 0; w < W; ++w) x[w] = valid ? __ldg(db + row * W + w) : 0u;
  const unsigned long long row_key = (unsigned long long)(idx_base + row);
  for (int j = 0; j < 32 && q0 + j < Q; ++j) {
    int d = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) d += __popc(x[w] ^ __ldg(q + (long long)(q0 + j) * W + w));
    cand_buf[(long long)(q0 + j) * cap + vrow] = valid ? (((unsigned long long)(unsigned)d << 40) | row_key) : ~0ull;
  }
}

// One CTA per query: sort the survivors (canonical keys: distance, row), keep the best k, tighten
// the threshold to the k-th distance (later rows may still tie with it at a smaller row index, so
// the test stays "<="); `final` writes keys_out.
__global__ void __launch_bounds__(CP_THREADS)
ham4_compact_kernel(unsigned long long* __restrict__ buf, int* __restrict__ cnt, int cap, int k, int K, int* __restrict__ tq,
                    int* __restrict__ overflow, int* __restrict__ tqmax, int final, unsigned long long* __restrict__ keys_out) {
  extern __shared__ unsigned long long s_key[];   // P = next pow2 >= min(count, cap)
  const int qi = blockIdx.x, tid = threadIdx.x;
  const int raw = cnt[qi];
  const int m = min(raw, cap);
  if (raw > cap && tid == 0) *overflow = 1;
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
  // rows past the table were stored as empty keys by the dense chunk: they sort last
  int keep = min(m, k);
  {
    int lo = 0, hi = keep;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_key[mid] != ~0ull) lo = mid + 1; else hi = mid;
    }
    keep = lo;
  }
  for (int i = tid; i < keep; i += CP_THREADS) mine[i] = s_key[i];
  if (tid == 0) {
    cnt[qi] = keep;
    const int t = (keep >= k) ? (int)(s_key[k - 1] >> 40) : K;
    tq[qi] = t;
    atomicMax(tqmax, t);                                         // T of the next chunk: its windows sit as tight as they can
  }
  if (final)
    for (int i = tid; i < k; i += CP_THREADS) keys_out[(size_t)qi * k + i] = (i < keep) ? s_key[i] : ~0ull;
}

struct HamTc4Plan {
  int K, qb, b_block, col_blocks, cols, cap, first_rows, growth, stages;
  size_t smem_bytes;
  size_t off_img, off_tq, off_cnt, off_flag, off_gcnt, off_list, off_buf, total;
  int list_cap, by_group;
};

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

int growth_for(int Q) {
  if (const char* e = getenv("SB_TC_GROWTH")) {               // tuning knob
    const int g = atoi(e);
    if (g >= 2 && g <= 64) return g;
  }
  return Q <= SMALL_Q ? GROWTH_SMALL_Q : GROWTH;
}

HamTc4Plan make_plan(int32_t W, int32_t Q, int32_t k) {
  HamTc4Plan p;
  p.K = 32 * W;
  // blocks of qb columns = 3 qb queries, qb a multiple of 32 (whole 32-column survivor groups) and <= 192: the
  // width with the least MMA time.  A kind::mxf4 MMA of N columns takes about 100 + 0.15 N cycles here (tile time
  // measured at N = 64 ... 192: 1700, 1810, 1970, 2060, 2080 cycles for 15 MMAs; 138 cycles at N = 256 in
  // tools/experiments/mxf4_probe2.cu), i.e. mostly a fixed cost per instruction: few wide blocks beat many narrow
  // ones even when they pad more.  4096 queries -> 8 x 192, 512 -> 1 x 192
  const int need = (Q + QPC - 1) / QPC;
  long long best = -1;
  p.qb = 32;
  for (int c = 32; c <= QB; c += 32) {
    const long long blocks = (need + c - 1) / c, cost = blocks * (c + 667);
    if (best < 0 || cost <= best) { best = cost; p.qb = c; }
  }
  if (const char* e = getenv("SB_TC4_QB")) {                    // tuning knob
    const int c = atoi(e);
    if (c >= 32 && c <= QB && c % 32 == 0) p.qb = c;
  }
  p.col_blocks = (need + p.qb - 1) / p.qb;
  p.cols = p.col_blocks * QPC * p.qb;                          // query slots
  p.b_block = (p.qb * 160 * QPC + 1023) / 1024 * 1024;         // 3 x (qb rows x 128 B) + 3 x B_syn (2 x qb x 16 B)
  p.cap = 4096;
  p.growth = growth_for(Q);
  while (p.cap < 4 * (p.growth + 1) * k) p.cap <<= 1;
  p.first_rows = Q <= SMALL_Q ? 512 : 256;                     // dense seed chunk: at least 8 k rows (its sort costs Q * rows)
  if (const char* e = getenv("SB_TC4_FIRST_ROWS")) {            // tuning knob
    const int r = atoi(e);
    if (r >= 32 && r <= 4096 && (r & (r - 1)) == 0) p.first_rows = r;
  }
  while (p.first_rows < 8 * k) p.first_rows <<= 1;
  if (p.first_rows > p.cap) p.first_rows = p.cap;
  p.stages = MAX_STAGES;
  p.smem_bytes = 1024 + (size_t)2 * p.b_block + A_SYN + (size_t)p.stages * A_GROUP + (2 * MAX_STAGES + 10) * 8;
  if (p.smem_bytes > 227 * 1024) {                             // the widest blocks leave room for two stages only
    p.stages = 2;
    p.smem_bytes -= (size_t)2 * A_GROUP;
  }
  size_t o = 0;
  p.off_img = o;  o += align256((size_t)p.col_blocks * p.b_block);
  p.off_tq = o;   o += align256((size_t)p.cols * sizeof(int));
  p.off_cnt = o;  o += align256((size_t)p.cols * sizeof(int));
  p.off_flag = o; o += 256;                                   // [0] overflow flag, [2], [3] T (by chunk parity)
  p.off_gcnt = o; o += align256((size_t)(p.cols / 32) * sizeof(int));   // re-check entries per 32-query group
  // entries per group list: a chunk yields ~(growth - 1) * (k + ties) survivors per query, 32 queries per group
  p.by_group = Q > SMALL_Q || k >= 32;                         // (many survivors per query: the group's queries in registers pay)
  if (p.by_group) {
    const long long need = 32ll * p.growth * (k + 32) * 4 / 3;
    p.list_cap = 8192;
    while (p.list_cap < need && p.list_cap < (1 << 16)) p.list_cap <<= 1;
    p.off_list = o; o += align256((size_t)(p.cols / 32) * p.list_cap * sizeof(unsigned long long));
  } else {
    p.list_cap = 1 << 21;                                      // one list for all (<= 1024) queries, k < 32
    p.off_list = o; o += align256((size_t)p.list_cap * sizeof(unsigned long long));
  }
  p.off_buf = o;  o += align256((size_t)p.cols * p.cap * sizeof(unsigned long long));
  p.total = o;
  return p;
}

long long gcd_ll(long long a, long long b) {
  while (b) { const long long t = a % b; a = b; b = t; }
  return a;
}

}  // namespace

extern "C" {

int sb_hamming_scan_tc4_supported(int64_t U, int32_t W, int32_t Q, int32_t k) {
  return U >= 1 && (W == 1 || W == 2 || W == 4 || W == 8) && Q >= 1 && k >= 1 && k <= 256 && U < (1ll << 38) && Q <= (1 << 21);   // (Q / 32 is a grid.y)
}

size_t sb_hamming_scan_tc4_workspace_bytes(int64_t U, int32_t W, int32_t Q, int32_t k) {
  if (!sb_hamming_scan_tc4_supported(U, W, Q, k)) return 0;
  return make_plan(W, Q, k).total;
}

// Same contract as sb_hamming_scan_tc (hamming_tc.cu), packed FP4 operands.
int sb_hamming_scan_tc4(const uint32_t* db, int64_t U, int32_t W, const uint32_t* q, int32_t Q, int32_t k, int64_t idx_base,
                        uint64_t* keys_out, int32_t* overflow_out, void* workspace, size_t workspace_bytes, void* stream) {
  SB_REQUIRE(db && q && keys_out && overflow_out, "sb_hamming_scan_tc4: NULL pointer");
  if (!sb_hamming_scan_tc4_supported(U, W, Q, k) || idx_base < 0 || idx_base + U >= (1ll << 40) ||
      (reinterpret_cast<uintptr_t>(db) & 15u)) {
    sb::set_error("sb_hamming_scan_tc4: needs W in {1, 2, 4, 8}, 1 <= k <= 256, U < 2^38, 16-byte aligned table");
    return SB_ERR_UNSUPPORTED;
  }
  const HamTc4Plan p = make_plan(W, Q, k);
  if (workspace == nullptr || workspace_bytes < p.total) {
    sb::set_error("sb_hamming_scan_tc4: workspace too small (%zu < %zu bytes)", workspace_bytes, p.total);
    return SB_ERR_WORKSPACE;
  }
  SB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "sb_hamming_scan_tc4: workspace must be 256-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  unsigned char* img = ws + p.off_img;
  int* tq = reinterpret_cast<int*>(ws + p.off_tq);
  int* cnt = reinterpret_cast<int*>(ws + p.off_cnt);
  int* flag = reinterpret_cast<int*>(ws + p.off_flag);
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(ws + p.off_buf);
  unsigned long long* list = reinterpret_cast<unsigned long long*>(ws + p.off_list);
  int* gcnt = reinterpret_cast<int*>(ws + p.off_gcnt);

  SB_CUDA_TRY(cudaMemsetAsync(img, 0, (size_t)p.col_blocks * p.b_block, st));
  {
    sb::ProfScope prof("ham_query_image_kernel", st);
    const long long items = (long long)Q * W;
    ham4_query_image_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(q, Q, W, p.qb, p.b_block, img);
    sb::count_launch();
    if (int rc = sb::check_launch("ham4_query_image_kernel")) return rc;
  }
  ham4_init_kernel<<<(p.cols + 255) / 256, 256, 0, st>>>(p.cols, p.K, tq, cnt, flag);
  sb::count_launch();
  if (int rc = sb::check_launch("ham4_init_kernel")) return rc;

  // visiting order: 32-row granules in a golden-ratio stride permutation (any prefix is spread evenly over the table)
  const long long NG = (U + GRAN - 1) / GRAN;
  long long P = 1;
  if (NG > 2) {
    P = (long long)(0.6180339887498949 * (double)NG);
    if (P < 1) P = 1;
    while (gcd_ll(P, NG) != 1) ++P;
    if (P >= NG) P = 1;
  }

  void (*kernel)(const HamTc4Params) =
      p.by_group ? (W == 8 ? ham_filter_fp4_kernel<8, true> : W == 4 ? ham_filter_fp4_kernel<4, true>
                    : W == 2 ? ham_filter_fp4_kernel<2, true> : ham_filter_fp4_kernel<1, true>)
                 : (W == 8 ? ham_filter_fp4_kernel<8, false> : W == 4 ? ham_filter_fp4_kernel<4, false>
                    : W == 2 ? ham_filter_fp4_kernel<2, false> : ham_filter_fp4_kernel<1, false>);
  SB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
  SB_CUDA_TRY(cudaFuncSetAttribute(ham4_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(p.cap * sizeof(unsigned long long))));
  const int sms = sb::sm_count();
  long long done = 0;                                         // granules
  for (int chunk = 0; done < NG; ++chunk) {
    int* const T_cur = flag + 2 + (chunk & 1);                  // written by the compaction behind the previous chunk
    int* const T_next = flag + 2 + ((chunk + 1) & 1);
    long long len = (done == 0) ? p.first_rows / GRAN : done * (p.growth - 1);
    if (len > NG - done) len = NG - done;
    const int dense = (done == 0) ? 1 : 0;
    if (dense) {
      // seed chunk: every pair kept (a few hundred rows): plain XOR / POPC
      const int n_vrows = (int)(len * GRAN);
      const dim3 grid((unsigned)((n_vrows + 127) / 128), (unsigned)((Q + 31) / 32));
      if (W == 8) ham4_seed_kernel<8><<<grid, 128, 0, st>>>(db, U, NG, P, n_vrows, q, Q, idx_base, buf, cnt, p.cap);
      else if (W == 4) ham4_seed_kernel<4><<<grid, 128, 0, st>>>(db, U, NG, P, n_vrows, q, Q, idx_base, buf, cnt, p.cap);
      else if (W == 2) ham4_seed_kernel<2><<<grid, 128, 0, st>>>(db, U, NG, P, n_vrows, q, Q, idx_base, buf, cnt, p.cap);
      else ham4_seed_kernel<1><<<grid, 128, 0, st>>>(db, U, NG, P, n_vrows, q, Q, idx_base, buf, cnt, p.cap);
      sb::count_launch();
      if (int rc = sb::check_launch("ham4_seed_kernel")) return rc;
    } else {
      ham4_threshold_image_kernel<<<(p.cols * 8 + 255) / 256, 256, 0, st>>>(Q, p.cols, p.K, p.qb, p.b_block, tq, T_cur, T_next, img, gcnt, flag);
      sb::count_launch();
      if (int rc = sb::check_launch("ham4_threshold_image_kernel")) return rc;
      HamTc4Params hp;
      hp.db = db; hp.U = U; hp.W = W; hp.vg0 = done; hp.vg1 = done + len; hp.NG = NG; hp.P = P;
      hp.col_blocks = p.col_blocks; hp.image = img; hp.tq = tq; hp.tqmax = T_cur; hp.recheck = list; hp.recheck_cnt = gcnt; hp.by_group = p.by_group;
      hp.recheck_cap = p.list_cap; hp.cand_buf = buf; hp.cand_cnt = cnt; hp.cap = p.cap;
      hp.idx_base = idx_base; hp.stages = p.stages; hp.qb = p.qb; hp.b_block = p.b_block;
      const long long n_tiles = (len + 3) / 4;
      const int gx = (int)(n_tiles < sms ? n_tiles : sms);
      int gy = sms / gx;
      if (gy < 1) gy = 1;
      if (gy > p.col_blocks) gy = p.col_blocks;
      hp.cb_per = (p.col_blocks + gy - 1) / gy;
      gy = (p.col_blocks + hp.cb_per - 1) / hp.cb_per;
      {
        sb::ProfScope prof("ham_filter_tc_kernel", st);
        kernel<<<dim3(gx, gy), THREADS, p.smem_bytes, st>>>(hp);
        sb::count_launch();
        if (int rc = sb::check_launch("ham_filter_fp4_kernel")) return rc;
      }
      sb::ProfScope prof("ham_recheck_kernel", st);
      if (!p.by_group) {
        const int blocks = 8 * sms;                             // 64 warps per SM: load-latency bound
        if (W == 8) ham4_recheck_flat_kernel<8><<<blocks, 256, 0, st>>>(db, q, Q, idx_base, list, gcnt, p.list_cap, tq, buf, cnt, p.cap, flag);
        else if (W == 4) ham4_recheck_flat_kernel<4><<<blocks, 256, 0, st>>>(db, q, Q, idx_base, list, gcnt, p.list_cap, tq, buf, cnt, p.cap, flag);
        else if (W == 2) ham4_recheck_flat_kernel<2><<<blocks, 256, 0, st>>>(db, q, Q, idx_base, list, gcnt, p.list_cap, tq, buf, cnt, p.cap, flag);
        else ham4_recheck_flat_kernel<1><<<blocks, 256, 0, st>>>(db, q, Q, idx_base, list, gcnt, p.list_cap, tq, buf, cnt, p.cap, flag);
        sb::count_launch();
        if (int rc = sb::check_launch("ham4_recheck_flat_kernel")) return rc;
      } else {
      // one CTA column per 32-query group, split so that the grid fills the GPU about four times over
      const int groups = (Q + 31) / 32;
      int split = (4 * sms + groups - 1) / groups;
      if (split < 1) split = 1;
      if (split > 32) split = 32;
      const dim3 blocks((unsigned)groups, (unsigned)split);
      if (W == 8) ham4_recheck_kernel<8><<<blocks, 256, 0, st>>>(db, q, Q, idx_base, list, gcnt, p.list_cap, tq, buf, cnt, p.cap, flag);
      else if (W == 4) ham4_recheck_kernel<4><<<blocks, 256, 0, st>>>(db, q, Q, idx_base, list, gcnt, p.list_cap, tq, buf, cnt, p.cap, flag);
      else if (W == 2) ham4_recheck_kernel<2><<<blocks, 256, 0, st>>>(db, q, Q, idx_base, list, gcnt, p.list_cap, tq, buf, cnt, p.cap, flag);
      else ham4_recheck_kernel<1><<<blocks, 256, 0, st>>>(db, q, Q, idx_base, list, gcnt, p.list_cap, tq, buf, cnt, p.cap, flag);
      sb::count_launch();
      if (int rc = sb::check_launch("ham4_recheck_kernel")) return rc;
      }
    }
    done += len;
    const int final = (done >= NG) ? 1 : 0;
    sb::ProfScope prof("ham_compact_kernel", st);
    ham4_compact_kernel<<<Q, CP_THREADS, p.cap * sizeof(unsigned long long), st>>>(buf, cnt, p.cap, k, p.K, tq, flag, T_next,
                                                                                   final, reinterpret_cast<unsigned long long*>(keys_out));
    sb::count_launch();
    if (int rc = sb::check_launch("ham4_compact_kernel")) return rc;
  }
  SB_CUDA_TRY(cudaMemcpyAsync(overflow_out, flag, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return SB_OK;
}

}  // extern "C"
