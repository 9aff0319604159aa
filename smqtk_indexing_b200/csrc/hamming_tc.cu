// Batched Hamming scan / top-k on the tensor cores (reference: LinearHashIndex._nn,
// smqtk_indexing/impls/hash_index/linear.py:232-240 + hamming_distance utils/metrics.py:155).
//
// Why: with thousands of queries per batch the XOR/POPC scan (hamming.cu) is bound by the integer
// issue pipes, not by HBM -- every table word is reused from registers by 4096 queries (27 TB/s of
// algorithmic bytes on a 6.6 TB/s HBM; profiles/r1_scan_full.md: ALU pipe 89 % busy).  The same
// count is a dot product: with bits mapped to +-1, dot(a, b) = K - 2 * popcount(a xor b).  FP8
// E4M3 holds +-1 exactly and every partial sum is an integer below 2048, exact even in an FP16
// accumulator, so tcgen05.mma kind::f8f6f4 computes the distances EXACTLY at 8192 MAC/clk/SM
// instead of 64 LOP3 + 16 POPC lanes/clk/SM.  The single-query / small-batch scan stays on
// hamming.cu, where the table really is streamed once per query and HBM is the bound.
//
// Data flow per CTA (persistent, one per SM):
//   * B = TWO blocks of 256 query codes, expanded once per batch to +-1 FP8 in the SWIZZLE_128B
//     K-major layout (ham_query_image_kernel), RESIDENT in shared memory for a whole pass over the
//     rows.
//   * A = 128 table rows per tile.  Producer warps read the PACKED codes (32 bytes per 256-bit
//     row: HBM traffic stays at U * b / 8 per pass) and expand them to FP8 straight into the
//     swizzled layout -- the table is never stored expanded.  The expansion is the most expensive
//     stage per tile (measured), so every expanded tile feeds both resident query blocks and the
//     producers get 8 of the 14 warps.
//   * The threshold rides in one extra K step: A_syn = 32 x (+1), B_syn = 32 slots that sum to
//     2 * tq - K + 1, so the accumulator is 2 * (tq - d) + 1 and "d <= tq" is a sign test (odd,
//     never zero).
//   * Accumulators: FP16 in TMEM, one 256-column buffer per resident query block, so block 1's MMAs
//     overlap block 0's epilogue.  The epilogue reads them packed (tcgen05.ld ... pack::16b, two
//     columns per register) and ANDs the sign bits of 32 columns at a time: no survivor, no work.
//   * Survivors are NOT decoded from registers: the lanes that saw one append (row, 32-query group)
//     to a global list and ham_recheck_kernel recomputes those few pairs with XOR/POPC.
// Thresholds tighten between chunks (ham_compact_kernel keeps the best k so far): a chunk is
// (GROWTH - 1) x the rows seen before it, and rows are visited in a golden-ratio permutation of
// 32-row granules, so every chunk is a uniform sample of the table whatever its order -- a chunk
// then yields ~(GROWTH - 1) * (k + ties) survivors per query for ANY data distribution (the table
// is sorted, near-duplicate clusters are contiguous).  A query whose buffer still overflows (a
// table of massively tied codes) raises a flag and the caller re-runs the batch on hamming.cu.
// Warp roles: 0-3 epilogue (TMEM lane quadrants), 4 MMA issue + TMEM alloc, 5 B loader,
// 6-13 producers (half a table row per thread and tile).
#include <cuda_fp16.h>

#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

using namespace tcptx;

namespace {

constexpr int TM = 128;                     // rows per tile (UMMA M)
constexpr int QB = 256;                     // query columns per block (UMMA N)
constexpr int A_GROUP = TM * 128;           // 16 KB: 128 rows x 128 FP8 (4 code words)
constexpr int B_GROUP = QB * 128;           // 32 KB
constexpr int B_SYN = 2 * QB * 16;          // 8 KB  (no-swizzle, 2 K chunks of 16 bytes)
constexpr int A_SYN = 2 * TM * 16;          // 4 KB
constexpr int MAX_STAGES = 4;
constexpr int NB = 2;                       // query blocks resident per pass: every expanded A tile feeds 2 x 256 queries
constexpr int GRAN = 32;                    // rows per granule of the visiting order
constexpr int EPI_WARPS = 4, MMA_WARP = 4, B_WARP = 5, PROD_WARP0 = 6, PROD_WARPS = 8;
constexpr int THREADS = (PROD_WARP0 + PROD_WARPS) * 32;   // 448
constexpr int GROWTH = 4;                   // rows of a chunk = 3 x the rows before it: ~3 (k + ties) survivors per query
constexpr int GROWTH_SMALL_Q = 8;           // few queries: every chunk's launch / fill / drain weighs more than the
constexpr int SMALL_Q = 1024;               // extra survivors -- 6 chunks of 7 (k + ties) instead of 9 of 3 (k + ties)
constexpr int CP_THREADS = 256;            // 7-8 compaction CTAs per SM: most queries hold a few hundred keys

struct HamTcParams {
  const uint32_t* db;          // u32[U][W]
  long long U;
  int W, G, ksteps;            // G = 128-byte K groups per row, ksteps = MMAs per group
  long long vg0, vg1;          // virtual granules of this chunk
  long long NG, P;             // physical granule = (virtual * P) mod NG
  int col_blocks, cb_per;      // query blocks in total / per blockIdx.y
  const unsigned char* image;  // per block: G x B_GROUP (SW128) then B_SYN
  const int* tq;               // thresholds (Hamming distance) per query column
  unsigned long long* recheck; // per 32-query group: rows with a survivor in the group, [group][recheck_cap]
  int* recheck_cnt;            // entries per group
  int recheck_cap;             // capacity of one group's list
  unsigned long long* cand_buf;
  int* cand_cnt;
  int cap;
  long long idx_base;
  int dense;                   // first chunk: every pair is kept -> key stored at buf[query][virtual row]
  int stages;
};

// 4 code bits -> 4 FP8 E4M3 bytes: bit = 0 -> +1.0 (0x38), bit = 1 -> -1.0 (0xB8); bit i -> byte i
__device__ __forceinline__ uint32_t expand_nibble(uint32_t x) { return ((x * 0x10204080u) & 0x80808080u) | 0x38383838u; }
__device__ __forceinline__ void expand_word(uint32_t w, uint4& lo, uint4& hi) {
  lo.x = expand_nibble(w & 15u);
  lo.y = expand_nibble((w >> 4) & 15u);
  lo.z = expand_nibble((w >> 8) & 15u);
  lo.w = expand_nibble((w >> 12) & 15u);
  hi.x = expand_nibble((w >> 16) & 15u);
  hi.y = expand_nibble((w >> 20) & 15u);
  hi.z = expand_nibble((w >> 24) & 15u);
  hi.w = expand_nibble(w >> 28);
}
// byte offset of 16-byte chunk c of row r inside a [rows x 128 B] SWIZZLE_128B K-major group
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
}

// D[tmem] (+)= A[smem] . B[smem]^T, kind::f8f6f4 (E4M3 x E4M3 -> F32), issued by one thread.
__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Survivors are rare (~GROWTH * k per query and chunk among millions of pairs), so the epilogue does
// not decode them from the accumulator registers (a divergent, register-indexed walk: ~400
// instructions per survivor, measured at 40 % of the kernel).  It only learns, per lane, WHICH
// 32-column group holds one and appends (row, group) to a global re-check list; ham_recheck_kernel
// then recomputes each listed row against the group's 32 queries with XOR/POPC from the packed
// codes -- one warp per entry, one query per lane, the whole GPU hiding the load latency that a
// re-check inside the epilogue would expose.
template <int G, int KSTEPS>
__global__ void __launch_bounds__(THREADS, 1) ham_filter_tc_kernel(const HamTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms: align by hand
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W = (G == 2) ? 8 : KSTEPS;                        // code words per row
  constexpr uint32_t a_stage = (uint32_t)G * A_GROUP;
  constexpr uint32_t b_block = (uint32_t)G * B_GROUP + B_SYN;      // one query block: data groups, then its B_syn
  // layout (offsets are multiples of 1024): [B block 0][B block 1][A_syn][A ring][barriers, queues, thresholds]
  unsigned char* s_b = smem;
  unsigned char* s_asyn = s_b + NB * b_block;
  unsigned char* s_a = s_asyn + A_SYN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_a + (size_t)p.stages * a_stage);
  const uint32_t a_full = smem_u32(bars), a_empty = a_full + MAX_STAGES * 8;
  const uint32_t acc_full = a_empty + MAX_STAGES * 8, acc_empty = acc_full + 16;
  const uint32_t b_full = acc_empty + 16, b_empty = b_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 6);
  int* s_tq = reinterpret_cast<int*>(bars + 2 * MAX_STAGES + 8);  // [4 warps][2 blocks x 256]: thresholds

  const long long n_gran = p.vg1 - p.vg0;
  const long long n_tiles = (n_gran + 3) / 4;
  const long long my_tiles = (n_tiles > (long long)blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int jb0 = blockIdx.y * p.cb_per;                         // cb_per is a multiple of NB
  const int jb1 = min(p.col_blocks, jb0 + p.cb_per);

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(a_full + s * 8, PROD_WARPS); mbar_init(a_empty + s * 8, 1); }
    for (int b = 0; b < NB; ++b) { mbar_init(acc_full + b * 8, 1); mbar_init(acc_empty + b * 8, EPI_WARPS); }
    mbar_init(b_full, 1);
    mbar_init(b_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // A_syn = +1.0 everywhere; codes narrower than 128 bits leave K chunks of the A ring untouched: zero them once
  for (int i = threadIdx.x; i < A_SYN / 4; i += THREADS) reinterpret_cast<uint32_t*>(s_asyn)[i] = 0x38383838u;
  if (KSTEPS < 4)
    for (int i = threadIdx.x; i < (int)(p.stages * a_stage / 16); i += THREADS)
      reinterpret_cast<uint4*>(s_a)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == B_WARP) {
    // =========================== B: NB resident query blocks ===========================
    if (lane == 0) {
      int it = 0;
      for (int jb = jb0; jb < jb1; jb += NB, ++it) {
        const int nb = min(NB, jb1 - jb);
        mbar_wait(b_empty, (it & 1) ^ 1);
        const unsigned char* src = p.image + (size_t)jb * b_block;   // blocks are contiguous in the image too
        const uint32_t bytes = (uint32_t)nb * b_block;
        mbar_expect_tx(b_full, bytes);
        for (uint32_t o = 0; o < bytes; o += 8192) tma_bulk_g2s(smem_u32(s_b + o), src + o, 8192, b_full);
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issue ===========================
    if (lane == 0) {
      // D = F16 (format 0: every partial sum is an integer below 2048, exact in half precision),
      // A = B = E4M3 (format 0), both K-major, N = 256, M = 128
      const uint32_t idesc = ((uint32_t)(QB >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      const uint64_t asyn_desc = umma_desc(smem_u32(s_asyn), TM * 16, 128);
      // descriptors differ only in the 14-bit start-address field (shared window < 256 KB: no carry out of it):
      // one 64-bit add per operand and MMA instead of rebuilding them
      const uint64_t a_desc0 = umma_desc_sw128(smem_u32(s_a));
      const uint64_t b_desc0 = umma_desc_sw128(smem_u32(s_b));
      const uint64_t bsyn_desc0 = umma_desc(smem_u32(s_b) + G * B_GROUP, QB * 16, 128);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      long long t[NB] = {0, 0};                                  // tiles issued per accumulator buffer
      for (int jb = jb0; jb < jb1; jb += NB, ++it) {
        const int nb = min(NB, jb1 - jb);
        mbar_wait(b_full, it & 1);
        tc_fence_after();
        for (long long i = 0; i < my_tiles; ++i) {
          mbar_wait(a_full + stage * 8, phase);
          const uint64_t a_desc = a_desc0 + (uint64_t)(((uint32_t)stage * a_stage) >> 4);
#pragma unroll
          for (int blk = 0; blk < NB; ++blk) {
            if (blk < nb) {
              // one A tile, NB query blocks: accumulator buffer = block, so block 1's MMAs overlap block 0's epilogue
              const uint32_t d_tmem = tmem_base + (uint32_t)(blk * QB);
              mbar_wait(acc_empty + blk * 8, (uint32_t)(t[blk] & 1) ^ 1);
              tc_fence_after();
              const uint64_t b_desc = b_desc0 + (uint64_t)((blk * b_block) >> 4);
#pragma unroll
              for (int g = 0; g < G; ++g)
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks)
                  umma_f8(d_tmem, a_desc + (uint64_t)((g * A_GROUP + ks * 32) >> 4), b_desc + (uint64_t)((g * B_GROUP + ks * 32) >> 4),
                          idesc, (g | ks) ? 1u : 0u);
              umma_f8(d_tmem, asyn_desc, bsyn_desc0 + (uint64_t)((blk * b_block) >> 4), idesc, 1u);
              umma_commit(acc_full + blk * 8);
              ++t[blk];
            }
          }
          umma_commit(a_empty + stage * 8);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(b_empty);                                   // B may be replaced once these MMAs retire
      }
    }
  } else if (warp >= PROD_WARP0) {
    // =========================== A: packed codes -> +-1 FP8, swizzled ===========================
    // thread -> (row r of the tile, half h of its words)
    constexpr int WH = (W >= 2) ? W / 2 : 1;                     // words per thread
    const int pt = threadIdx.x - PROD_WARP0 * 32;
    const int r = pt & (TM - 1), h = pt >> 7;
    const bool worker = (W >= 2) || h == 0;
    int stage = 0;
    uint32_t phase = 0;
    // PF tiles of packed words are in flight per thread: with few queries (one pass per chunk) every tile comes
    // straight from HBM, one round trip (~1.5 us under load) is longer than a tile's MMAs (1.2 us), and a single
    // tile of look-ahead made the LOAD LATENCY the pace: 3170 cycles per tile at Q = 512 vs 2300 at Q = 4096
    constexpr int PF = 4;
    uint32_t cw[PF][WH];
    bool cv[PF];
    // granule of tile i for this thread: vg = vg_first + i * 4 * gridDim.x, physical pg = vg * P mod NG, kept incrementally
    const long long vg_first = p.vg0 + (long long)blockIdx.x * 4 + (r >> 5);
    const long long pg_first = (long long)(((unsigned long long)vg_first * (unsigned long long)p.P) % (unsigned long long)p.NG);
    const long long pg_step = (long long)(((unsigned long long)(4 * gridDim.x) * (unsigned long long)p.P) % (unsigned long long)p.NG);
    long long vg_ld = vg_first, pg_ld = pg_first;
    auto load = [&](uint32_t (&w)[WH]) {                        // loads the NEXT tile in sequence
      const bool in_chunk = vg_ld < p.vg1;
      const long long row = pg_ld * GRAN + lane;
      const bool valid = in_chunk && row < p.U;
      vg_ld += 4 * gridDim.x;
      pg_ld += pg_step;
      if (pg_ld >= p.NG) pg_ld -= p.NG;
#pragma unroll
      for (int j = 0; j < WH; ++j) w[j] = 0u;
      if (valid && worker) {
        const uint32_t* src = p.db + row * W + h * WH;
        if (WH == 4) {
          const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(src));
          w[0] = x0.x; w[1] = x0.y; w[2 % WH] = x0.z; w[3 % WH] = x0.w;
        } else if (WH == 2) {
          const uint2 x0 = __ldg(reinterpret_cast<const uint2*>(src));
          w[0] = x0.x; w[1 % WH] = x0.y;
        } else {
          w[0] = __ldg(src);
        }
      }
      return valid;
    };
    for (int jb = jb0; jb < jb1; jb += NB) {
      vg_ld = vg_first;
      pg_ld = pg_first;
#pragma unroll
      for (int d = 0; d < PF; ++d) cv[d] = (d < my_tiles) ? load(cw[d]) : false;
      for (long long i0 = 0; i0 < my_tiles; i0 += PF) {
#pragma unroll
        for (int d = 0; d < PF; ++d) {
          if (i0 + d < my_tiles) {
            mbar_wait(a_empty + stage * 8, phase ^ 1);
            unsigned char* dst = s_a + (size_t)stage * a_stage;
            if (worker) {
#pragma unroll
              for (int j = 0; j < WH; ++j) {
                const int wj = h * WH + j;                         // word of the row
                uint4 lo, hi;
                expand_word(cw[d][j], lo, hi);
                if (!cv[d]) { lo = make_uint4(0u, 0u, 0u, 0u); hi = lo; }   // rows past the table: all-zero operand
                unsigned char* grp = dst + (wj >> 2) * A_GROUP;
                *reinterpret_cast<uint4*>(grp + sw128_off(r, 2 * (wj & 3))) = lo;
                *reinterpret_cast<uint4*>(grp + sw128_off(r, 2 * (wj & 3) + 1)) = hi;
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + stage * 8);
            cv[d] = (i0 + d + PF < my_tiles) ? load(cw[d]) : false;   // refill this slot with the tile PF ahead
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp < EPI_WARPS) {
    // =========================== epilogue: sign test, survivors queued ===========================
    const int ew = warp;                                         // TMEM lane quadrant = granule of the tile
    uint32_t par = 0u;                                           // bit blk: parity of accumulator buffer blk's next "full"
    int* my_tq = s_tq + warp * (NB * QB);
    const long long vg_first = p.vg0 + (long long)blockIdx.x * 4 + ew;
    const long long pg_first = (long long)(((unsigned long long)vg_first * (unsigned long long)p.P) % (unsigned long long)p.NG);
    const long long pg_step = (long long)(((unsigned long long)(4 * gridDim.x) * (unsigned long long)p.P) % (unsigned long long)p.NG);
    for (int jb = jb0; jb < jb1; jb += NB) {
      const int nb = min(NB, jb1 - jb);
      __syncwarp();
      for (int c = lane; c < nb * QB; c += 32) my_tq[c] = __ldcg(p.tq + (long long)jb * QB + c);
      __syncwarp();
      long long vg = vg_first, pg = pg_first;
      for (long long i = 0; i < my_tiles; ++i) {
        const long long vt = blockIdx.x + i * gridDim.x;
        const long long row = pg * GRAN + lane;
        const bool rvalid = (vg < p.vg1) && (row < p.U);
        vg += 4 * gridDim.x;
        pg += pg_step;
        if (pg >= p.NG) pg -= p.NG;
        const unsigned long long row_key = (unsigned long long)(p.idx_base + row);
#pragma unroll 1
        for (int blk = 0; blk < nb; ++blk) {
          const int q0 = (jb + blk) * QB;
          const int* tq_blk = my_tq + blk * QB;
          mbar_wait(acc_full + blk * 8, (par >> blk) & 1u);
          par ^= 1u << blk;
          tc_fence_after();
          const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(blk * QB);
          unsigned hit[QB / 32];                                     // lanes with a survivor in each 32-column group
#pragma unroll
          for (int c128 = 0; c128 < QB; c128 += 128) {
            uint32_t va[32], vb[32];                               // 2 x 64 columns of packed FP16 accumulators
            tmem_ld32_pack16_nowait(tbase + (uint32_t)c128, va);
            tmem_ld32_pack16_nowait(tbase + (uint32_t)(c128 + 64), vb);
            tmem_ld_wait();
            if (p.dense) {
              // seed chunk: buf[query][virtual row] = key for every pair (virtual row < cap by construction)
              const long long vrow = vt * TM + ew * 32 + lane;
#pragma unroll
              for (int j = 0; j < 64; ++j) {
                const uint32_t reg = (j < 32) ? va[j & 31] : vb[j & 31];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  const int col = c128 + 2 * j + hh;
                  const int d = tq_blk[col] - (int)(__half2float(__ushort_as_half((unsigned short)(reg >> (16 * hh)))) * 0.5f);
                  p.cand_buf[(long long)(q0 + col) * p.cap + vrow] = rvalid ? (((unsigned long long)(unsigned)d << 40) | row_key) : ~0ull;
                }
              }
            } else {
              // the common case (no survivor among 64 pairs) is an AND tree over the packed sign bits
              uint32_t a0 = 0xffffffffu, a1 = 0xffffffffu, b0 = 0xffffffffu, b1 = 0xffffffffu;
#pragma unroll
              for (int j = 0; j < 16; ++j) { a0 &= va[j]; a1 &= va[16 + j]; b0 &= vb[j]; b1 &= vb[16 + j]; }
              hit[c128 / 32] = __ballot_sync(0xffffffffu, rvalid && (a0 & 0x80008000u) != 0x80008000u);
              hit[c128 / 32 + 1] = __ballot_sync(0xffffffffu, rvalid && (a1 & 0x80008000u) != 0x80008000u);
              hit[c128 / 32 + 2] = __ballot_sync(0xffffffffu, rvalid && (b0 & 0x80008000u) != 0x80008000u);
              hit[c128 / 32 + 3] = __ballot_sync(0xffffffffu, rvalid && (b1 & 0x80008000u) != 0x80008000u);
            }
          }
          // the accumulator buffer is free: let the next tile's MMAs start, THEN pay for the appends
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty + blk * 8);
          if (!p.dense) {
            unsigned any = 0u;
#pragma unroll
            for (int g32 = 0; g32 < QB / 32; ++g32) any |= hit[g32];
            if (any) {                                                 // rare: survivors in this (warp, tile, block)
              // one slot reservation per group WITH survivors, all issued at once: lane g reserves for group g
              // (serial atomics would cost a global round trip each while the next tile's accumulator waits)
              int mine_cnt = 0;
#pragma unroll
              for (int g32 = 0; g32 < QB / 32; ++g32) mine_cnt = (lane == g32) ? __popc(hit[g32]) : mine_cnt;
              int mine_base = 0;
              if (lane < QB / 32 && mine_cnt) mine_base = atomicAdd(p.recheck_cnt + (q0 >> 5) + lane, mine_cnt);
#pragma unroll
              for (int g32 = 0; g32 < QB / 32; ++g32) {
                const unsigned m = hit[g32];
                const int base = __shfl_sync(0xffffffffu, mine_base, g32);
                if ((m >> lane) & 1u) {
                  const int slot = base + __popc(m & ((1u << lane) - 1u));
                  if (slot < p.recheck_cap) p.recheck[(size_t)((q0 >> 5) + g32) * p.recheck_cap + slot] = (unsigned long long)row;
                }
              }
            }
          }
          __syncwarp();
        }
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

// Re-check, one CTA column per 32-query group (blockIdx.x = group, blockIdx.y splits the group's list): every lane
// keeps ITS query of the group in registers for the whole kernel, so an entry costs one 32-byte row read instead
// of the row plus the group's 32 query codes (1 KB from L2 per entry with the unordered global list of round 1).
// A warp takes 32 entries at a time: lane l loads entry l's row, then the 32 rows are broadcast one after the
// other with shuffles and every lane tests its own query -- 32 independent loads in flight per warp.
template <int W>
__global__ void __launch_bounds__(256)
ham_recheck_kernel(const uint32_t* __restrict__ db, const uint32_t* __restrict__ qcodes, int Q, long long idx_base,
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

// Query codes -> per-block image: G groups of [256 rows x 128 B] in the SWIZZLE_128B K-major order,
// same bit -> byte expansion as the table rows; columns past Q and K chunks past the code are zero
// (the image is cleared first).  One thread per (column, word).
__global__ void ham_query_image_kernel(const uint32_t* __restrict__ q, int Q, int W, int G, unsigned char* __restrict__ img) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Q * W) return;
  const int col = (int)(i / W), j = (int)(i % W);
  const int jb = col / QB, n = col % QB;
  uint4 lo, hi;
  expand_word(q[(long long)col * W + j], lo, hi);
  unsigned char* grp = img + (size_t)jb * ((size_t)G * B_GROUP + B_SYN) + (size_t)(j >> 2) * B_GROUP;
  *reinterpret_cast<uint4*>(grp + sw128_off(n, 2 * (j & 3))) = lo;
  *reinterpret_cast<uint4*>(grp + sw128_off(n, 2 * (j & 3) + 1)) = hi;
}

// B_syn of one column: 32 E4M3 slots that sum to s = 2 * tq - K + 1 (|s| <= 16 * 32); padding columns get -512
// (their data bytes are zero: the accumulator is negative).
__device__ __forceinline__ void write_threshold_slots(unsigned char* __restrict__ img, int col, int G, int s) {
  const int jb = col / QB, n = col % QB;
  const uint32_t sign = s < 0 ? 0x80u : 0u;
  const int mag = s < 0 ? -s : s;
  const int n16 = mag >> 4, rem = mag & 15;
  // E4M3 codes of the integers 0..15 (16 = 0x58)
  const unsigned char e4m3[16] = {0x00, 0x38, 0x40, 0x44, 0x48, 0x4A, 0x4C, 0x4E, 0x50, 0x51, 0x52, 0x53, 0x54, 0x55, 0x56, 0x57};
  uint32_t words[8];
#pragma unroll
  for (int wd = 0; wd < 8; ++wd) {
    uint32_t v = 0u;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int slot = wd * 4 + b;
      uint32_t byte = 0u;
      if (slot < n16) byte = 0x58u | sign;
      else if (slot == n16 && rem) byte = (uint32_t)e4m3[rem] | sign;
      v |= byte << (8 * b);
    }
    words[wd] = v;
  }
  unsigned char* base = img + (size_t)jb * ((size_t)G * B_GROUP + B_SYN) + (size_t)G * B_GROUP;
  *reinterpret_cast<uint4*>(base + n * 16) = make_uint4(words[0], words[1], words[2], words[3]);
  *reinterpret_cast<uint4*>(base + QB * 16 + n * 16) = make_uint4(words[4], words[5], words[6], words[7]);
}

// Start of a batch: no threshold yet (tq = K: every pair passes), empty candidate lists, the threshold K step of
// every column (padding columns: always negative), empty re-check list.
__global__ void ham_init_kernel(int Q, int cols, int G, int K, int* __restrict__ tq, int* __restrict__ cnt,
                                int* __restrict__ overflow, unsigned char* __restrict__ img, int* __restrict__ list_cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *overflow = 0;
  if (i >= cols) return;
  if ((i & 31) == 0) list_cnt[i >> 5] = 0;                       // one re-check list per 32-query group
  tq[i] = K;
  cnt[i] = 0;
  write_threshold_slots(img, i, G, i < Q ? K + 1 : -512);
}

// One CTA per query: sort the survivors (canonical keys: distance, row), keep the best k, tighten
// the threshold to the k-th distance (later rows may still tie with it at a smaller row index, so
// the test stays "<="); `final` writes keys_out.
__global__ void __launch_bounds__(CP_THREADS)
ham_compact_kernel(unsigned long long* __restrict__ buf, int* __restrict__ cnt, int cap, int k, int K, int* __restrict__ tq,
                   int* __restrict__ overflow, int final, unsigned long long* __restrict__ keys_out, int dense_count,
                   unsigned char* __restrict__ img, int G, int* __restrict__ list_cnt) {
  extern __shared__ unsigned long long s_key[];   // P = next pow2 >= min(count, cap)
  const int qi = blockIdx.x, tid = threadIdx.x;
  const int raw = dense_count >= 0 ? dense_count : cnt[qi];      // the dense seed chunk filled buf[query][0 .. rows)
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
    // the next chunk's threshold K step for this query, and an empty re-check list for its group (this kernel runs
    // after the re-check of the chunk: one launch per chunk less than a separate threshold-image kernel)
    if (!final) write_threshold_slots(img, qi, G, 2 * min(t, K) - K + 1);
    if ((qi & 31) == 0) list_cnt[qi >> 5] = 0;
  }
  if (final)
    for (int i = tid; i < k; i += CP_THREADS) keys_out[(size_t)qi * k + i] = (i < keep) ? s_key[i] : ~0ull;
}

struct HamTcPlan {
  int G, ksteps, K, col_blocks, cols, cap, first_rows, stages, growth;
  size_t block_bytes, smem_bytes;
  size_t off_img, off_tq, off_cnt, off_flag, off_gcnt, off_list, off_buf, total;
  int list_cap;
};

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

int growth_for(int Q) {
  if (const char* e = getenv("SB_TC_GROWTH")) {               // tuning knob
    const int g = atoi(e);
    if (g >= 2 && g <= 64) return g;
  }
  return Q <= SMALL_Q ? GROWTH_SMALL_Q : GROWTH;
}

HamTcPlan make_plan(int32_t W, int32_t Q, int32_t k) {
  HamTcPlan p;
  p.G = W <= 4 ? 1 : W / 4;
  p.ksteps = W < 4 ? W : 4;
  p.K = 32 * W;
  p.col_blocks = (Q + QB - 1) / QB;
  p.cols = p.col_blocks * QB;
  p.cap = 4096;
  p.growth = growth_for(Q);
  while (p.cap < 4 * (p.growth + 1) * k) p.cap <<= 1;
  p.first_rows = Q <= SMALL_Q ? 1024 : 256;                    // dense seed chunk: at least 8 k rows (its sort costs Q * rows)
  while (p.first_rows < 8 * k) p.first_rows <<= 1;
  if (p.first_rows > p.cap) p.first_rows = p.cap;
  p.stages = p.G == 2 ? 2 : MAX_STAGES;
  p.block_bytes = (size_t)p.G * B_GROUP + B_SYN;
  p.smem_bytes = 1024 + NB * p.block_bytes + A_SYN + (size_t)p.stages * p.G * A_GROUP + (2 * MAX_STAGES + 8) * 8 +
                 EPI_WARPS * NB * QB * sizeof(int);
  size_t o = 0;
  p.off_img = o;  o += align256((size_t)p.col_blocks * p.block_bytes);
  p.off_tq = o;   o += align256((size_t)p.cols * sizeof(int));
  p.off_cnt = o;  o += align256((size_t)p.cols * sizeof(int));
  p.off_flag = o; o += 256;                                   // [0] overflow flag
  p.off_gcnt = o; o += align256((size_t)(p.cols / 32) * sizeof(int));   // re-check entries per 32-query group
  // entries per group list: a chunk yields ~(growth - 1) * (k + ties) survivors per query, 32 queries per group
  {
    const long long need = 32ll * p.growth * (k + 32) * 4 / 3;
    p.list_cap = 8192;
    while (p.list_cap < need && p.list_cap < (1 << 16)) p.list_cap <<= 1;
  }
  p.off_list = o; o += align256((size_t)(p.cols / 32) * p.list_cap * sizeof(unsigned long long));
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

int sb_hamming_scan_tc_supported(int64_t U, int32_t W, int32_t Q, int32_t k) {
  return U >= 1 && (W == 1 || W == 2 || W == 4 || W == 8) && Q >= 1 && k >= 1 && k <= 256 && U < (1ll << 38);
}

size_t sb_hamming_scan_tc_workspace_bytes(int64_t U, int32_t W, int32_t Q, int32_t k) {
  if (!sb_hamming_scan_tc_supported(U, W, Q, k)) return 0;
  return make_plan(W, Q, k).total;
}

// Same contract as sb_hamming_scan; *overflow_out (device int) is set to 1 when a candidate buffer
// overflowed -- the keys are then NOT trustworthy and the caller must re-run sb_hamming_scan.
int sb_hamming_scan_tc(const uint32_t* db, int64_t U, int32_t W, const uint32_t* q, int32_t Q, int32_t k, int64_t idx_base,
                       uint64_t* keys_out, int32_t* overflow_out, void* workspace, size_t workspace_bytes, void* stream) {
  SB_REQUIRE(db && q && keys_out && overflow_out, "sb_hamming_scan_tc: NULL pointer");
  if (!sb_hamming_scan_tc_supported(U, W, Q, k) || idx_base < 0 || idx_base + U >= (1ll << 40) ||
      (reinterpret_cast<uintptr_t>(db) & 15u)) {
    sb::set_error("sb_hamming_scan_tc: needs W in {1, 2, 4, 8}, 1 <= k <= 256, U < 2^38, 16-byte aligned table");
    return SB_ERR_UNSUPPORTED;
  }
  const HamTcPlan p = make_plan(W, Q, k);
  if (workspace == nullptr || workspace_bytes < p.total) {
    sb::set_error("sb_hamming_scan_tc: workspace too small (%zu < %zu bytes)", workspace_bytes, p.total);
    return SB_ERR_WORKSPACE;
  }
  SB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "sb_hamming_scan_tc: workspace must be 256-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  unsigned char* img = ws + p.off_img;
  int* tq = reinterpret_cast<int*>(ws + p.off_tq);
  int* cnt = reinterpret_cast<int*>(ws + p.off_cnt);
  int* flag = reinterpret_cast<int*>(ws + p.off_flag);
  int* gcnt = reinterpret_cast<int*>(ws + p.off_gcnt);
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(ws + p.off_buf);
  unsigned long long* list = reinterpret_cast<unsigned long long*>(ws + p.off_list);

  SB_CUDA_TRY(cudaMemsetAsync(img, 0, (size_t)p.col_blocks * p.block_bytes, st));
  {
    sb::ProfScope prof("ham_query_image_kernel", st);
    const long long items = (long long)Q * W;
    ham_query_image_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(q, Q, W, p.G, img);
    sb::count_launch();
    if (int rc = sb::check_launch("ham_query_image_kernel")) return rc;
  }
  ham_init_kernel<<<(p.cols + 255) / 256, 256, 0, st>>>(Q, p.cols, p.G, p.K, tq, cnt, flag, img, gcnt);
  sb::count_launch();
  if (int rc = sb::check_launch("ham_init_kernel")) return rc;

  // visiting order: 32-row granules in a golden-ratio stride permutation (any prefix is spread evenly over the table)
  const long long NG = (U + GRAN - 1) / GRAN;
  long long P = 1;
  if (NG > 2) {
    P = (long long)(0.6180339887498949 * (double)NG);
    if (P < 1) P = 1;
    while (gcd_ll(P, NG) != 1) ++P;
    if (P >= NG) P = 1;
  }

  void (*kernel)(const HamTcParams) = W == 8   ? ham_filter_tc_kernel<2, 4>
                                      : W == 4 ? ham_filter_tc_kernel<1, 4>
                                      : W == 2 ? ham_filter_tc_kernel<1, 2>
                                               : ham_filter_tc_kernel<1, 1>;
  SB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
  SB_CUDA_TRY(cudaFuncSetAttribute(ham_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(p.cap * sizeof(unsigned long long))));
  const int sms = sb::sm_count();
  long long done = 0;                                         // granules
  while (done < NG) {
    long long len = (done == 0) ? p.first_rows / GRAN : done * (p.growth - 1);
    if (len > NG - done) len = NG - done;
    const int dense = (done == 0) ? 1 : 0;
    HamTcParams hp;
    hp.db = db; hp.U = U; hp.W = W; hp.G = p.G; hp.ksteps = p.ksteps; hp.vg0 = done; hp.vg1 = done + len; hp.NG = NG; hp.P = P;
    hp.col_blocks = p.col_blocks; hp.image = img; hp.tq = tq; hp.recheck = list; hp.recheck_cnt = gcnt; hp.recheck_cap = p.list_cap; hp.cand_buf = buf; hp.cand_cnt = cnt; hp.cap = p.cap;
    hp.idx_base = idx_base; hp.dense = dense; hp.stages = p.stages;
    const long long n_tiles = (len + 3) / 4;
    const int gx = (int)(n_tiles < sms ? n_tiles : sms);
    int gy = sms / gx;
    if (gy < 1) gy = 1;
    if (gy > p.col_blocks) gy = p.col_blocks;
    hp.cb_per = (p.col_blocks + gy - 1) / gy;
    hp.cb_per = (hp.cb_per + NB - 1) / NB * NB;                 // whole groups of resident query blocks
    gy = (p.col_blocks + hp.cb_per - 1) / hp.cb_per;
    {
      sb::ProfScope prof("ham_filter_tc_kernel", st);
      kernel<<<dim3(gx, gy), THREADS, p.smem_bytes, st>>>(hp);
      sb::count_launch();
      if (int rc = sb::check_launch("ham_filter_tc_kernel")) return rc;
    }
    if (!dense) {
      sb::ProfScope prof("ham_recheck_kernel", st);
      // one CTA column per 32-query group, split so that the grid fills the GPU about four times over
      const int groups = (Q + 31) / 32;
      int split = (4 * sms + groups - 1) / groups;
      if (split < 1) split = 1;
      if (split > 32) split = 32;
      const dim3 blocks((unsigned)groups, (unsigned)split);
      if (W == 8) ham_recheck_kernel<8><<<blocks, 256, 0, st>>>(db, q, Q, idx_base, list, gcnt, p.list_cap, tq, buf, cnt, p.cap, flag);
      else if (W == 4) ham_recheck_kernel<4><<<blocks, 256, 0, st>>>(db, q, Q, idx_base, list, gcnt, p.list_cap, tq, buf, cnt, p.cap, flag);
      else if (W == 2) ham_recheck_kernel<2><<<blocks, 256, 0, st>>>(db, q, Q, idx_base, list, gcnt, p.list_cap, tq, buf, cnt, p.cap, flag);
      else ham_recheck_kernel<1><<<blocks, 256, 0, st>>>(db, q, Q, idx_base, list, gcnt, p.list_cap, tq, buf, cnt, p.cap, flag);
      sb::count_launch();
      if (int rc = sb::check_launch("ham_recheck_kernel")) return rc;
    }
    done += len;
    const int final = (done >= NG) ? 1 : 0;
    sb::ProfScope prof("ham_compact_kernel", st);
    ham_compact_kernel<<<Q, CP_THREADS, p.cap * sizeof(unsigned long long), st>>>(
        buf, cnt, p.cap, k, p.K, tq, flag, final, reinterpret_cast<unsigned long long*>(keys_out),
        dense ? (int)(len * GRAN) : -1, img, p.G, gcnt);
    sb::count_launch();
    if (int rc = sb::check_launch("ham_compact_kernel")) return rc;
  }
  SB_CUDA_TRY(cudaMemcpyAsync(overflow_out, flag, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return SB_OK;
}

}  // extern "C"
