// Stage 2 of the LSH path: brute-force Hamming scan with per-warp top-k lists,
// followed by a list merge.  Replaces LinearHashIndex._nn
// (reference: smqtk_indexing/impls/hash_index/linear.py:232-244, distance
// function smqtk_indexing/utils/metrics.py:155).
//
// Mapping (sm_100a):
//   grid  = (code chunks, query tiles);  CTA = 8 warps.
//   A warp owns 32*C consecutive-by-lane codes in REGISTERS (coalesced 128-bit
//   loads straight from HBM: lane l reads row tile+c*32+l), then walks the CTA's
//   query tile, which sits in shared memory (staged once by a TMA bulk copy) and
//   is read with warp-broadcast LDS.128.  Every loaded code word is therefore
//   reused QT times from registers; HBM sees the table once per query tile.
//   Per pair: XOR (LOP3) + carry-save-adder tree (LOP3) + POPC + compare with the
//   query's running threshold tau.  tau is shared by the CTA's warps (smem,
//   atomicMin); candidates that pass take a warp-cooperative slow path (ballot /
//   shuffle) that inserts them into the warp's sorted k-list for that query.
//   After warm-up the slow path fires ~k*ln(n/k) times per list.
//   The CTA folds its 8 warp lists into one, so the result is P = chunks sorted
//   lists per query -> merge kernel (distance-histogram select) -> top-k keys.
//
// Exactness: integer arithmetic only; keys (distance<<40 | row) are unique, the
// merge is a pure selection, so the result is independent of chunking.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace {

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int TAU_INIT = 0x7ffffffe;   // passes everything except invalid lanes
constexpr int D_INVALID = 0x7fffffff;

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// a * m + b as an IMAD: `m` is an opaque register holding 1 (a kernel argument), so the
// add is issued to the FMA pipe, which the scan otherwise leaves idle, instead of the
// ALU pipe that carries the LOP3 stream.
__device__ __forceinline__ int fma_add(int a, int b, int m) { return a * m + b; }

// popcount of 8 xor-ed words.
//   MODE 0: 8 POPC.
//   MODE 1: three carry-save adders fold the 8 words into 2 weight-1 and 3 weight-2
//           words -> 5 POPC + 6 LOP3.
//   MODE 2: a fourth adder folds the weight-2 words -> 4 POPC + 8 LOP3, sums on the FMA
//           pipe.  POPC runs on the XU pipe at 16 lanes/clk/SM, LOP3 on the ALU pipe at
//           64: per pair MODE 2 is 32 XU cycles vs 2*(8 XOR + 8 LOP3 + 1 ISETP) = 34
//           ALU cycles -- the two pipes finish together.
template <int MODE>
__device__ __forceinline__ int popc8(const uint32_t* x, int one) {
  if (MODE == 0) {
    return (__popc(x[0]) + __popc(x[1]) + __popc(x[2])) + (__popc(x[3]) + __popc(x[4]) + __popc(x[5])) +
           (__popc(x[6]) + __popc(x[7]));
  } else {
    uint32_t s1 = xor3(x[0], x[1], x[2]), c1 = maj3(x[0], x[1], x[2]);
    uint32_t s2 = xor3(x[3], x[4], x[5]), c2 = maj3(x[3], x[4], x[5]);
    uint32_t s3 = xor3(s1, s2, x[6]), c3 = maj3(s1, s2, x[6]);
    if (MODE == 1) {
      int ones = __popc(s3) + __popc(x[7]);
      int twos = __popc(c1) + __popc(c2) + __popc(c3);
      return ones + 2 * twos;
    } else {
      uint32_t s4 = xor3(c1, c2, c3), c4 = maj3(c1, c2, c3);     // weight 2, weight 4
      int d = fma_add(__popc(s3), __popc(x[7]), one);
      d = __popc(s4) * 2 + d;
      return __popc(c4) * 4 + d;
    }
  }
}

template <int W, int MODE>
__device__ __forceinline__ int hamming(const uint32_t (&c)[W], const uint32_t (&q)[W], int one) {
  uint32_t x[W];
#pragma unroll
  for (int i = 0; i < W; ++i) x[i] = c[i] ^ q[i];
  if constexpr (W == 1) {
    return __popc(x[0]);
  } else if constexpr (W == 2) {
    return __popc(x[0]) + __popc(x[1]);
  } else if constexpr (W == 4) {
    if (MODE == 0) return __popc(x[0]) + __popc(x[1]) + __popc(x[2]) + __popc(x[3]);
    uint32_t s = xor3(x[0], x[1], x[2]), cy = maj3(x[0], x[1], x[2]);
    return __popc(cy) * 2 + fma_add(__popc(s), __popc(x[3]), one);
  } else {
    int d = 0;
#pragma unroll
    for (int g = 0; g + 8 <= W; g += 8) d = fma_add(popc8<MODE>(x + g, one), d, one);
    return d;
  }
}

template <int W>
__device__ __forceinline__ void load_words(uint32_t (&dst)[W], const uint32_t* __restrict__ src) {
  if constexpr (W == 1) {
    dst[0] = __ldg(src);
  } else if constexpr (W == 2) {
    uint2 v = __ldg(reinterpret_cast<const uint2*>(src));
    dst[0] = v.x; dst[1] = v.y;
  } else {
#pragma unroll
    for (int i = 0; i < W / 4; ++i) {
      uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
      dst[4 * i + 0] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
    }
  }
}

template <int W>
__device__ __forceinline__ void load_words_shared(uint32_t (&dst)[W], const uint32_t* src) {
  if constexpr (W == 1) {
    dst[0] = src[0];
  } else if constexpr (W == 2) {
    uint2 v = *reinterpret_cast<const uint2*>(src);
    dst[0] = v.x; dst[1] = v.y;
  } else {
#pragma unroll
    for (int i = 0; i < W / 4; ++i) {
      uint4 v = reinterpret_cast<const uint4*>(src)[i];
      dst[4 * i + 0] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
    }
  }
}

// Insert `key` into the ascending k-list `list` (shared memory, owned by this
// warp), dropping the largest element.  All 32 lanes call this together.
__device__ __forceinline__ void warp_insert(uint64_t* list, int k, uint64_t key, int lane) {
  if (key >= list[k - 1]) return;  // warp-uniform
  int pos = 0;
  for (int base = 0; base < k; base += 32) {
    int i = base + lane;
    bool lt = (i < k) && (list[i] < key);
    pos += __popc(__ballot_sync(sb::FULL_MASK, lt));
  }
  for (int base = ((k - 1) / 32) * 32; base >= 0 && base + 31 >= pos; base -= 32) {
    int i = base + lane;
    bool mv = (i >= pos) && (i < k - 1);
    uint64_t v = mv ? list[i] : 0ull;
    __syncwarp();
    if (mv) list[i + 1] = v;
    __syncwarp();
  }
  if (lane == 0) list[pos] = key;
  __syncwarp();
}

// ---- TMA bulk copy (1-D) of the query tile into shared memory ---------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- fold the CTA's 8 warp lists into one list per query: part[q][chunk][k] ----
// Keys are unique, so the rank of a key among all 8*k candidates is the sum of its
// lower bounds in the 8 sorted lists; ranks < k are written straight to HBM.
__device__ __forceinline__ void fold_lists(const uint64_t* lists, int QT, int nq, int k, uint64_t* __restrict__ part,
                                           int P, int qt0, int tid) {
  __syncthreads();
  for (int i = tid; i < nq * k; i += THREADS) {
    const int qi = i / k, j = i - qi * k;
    part[((size_t)(qt0 + qi) * P + blockIdx.x) * k + j] = SB_KEY_EMPTY;
  }
  __syncthreads();
  const int per_q = WARPS * k;
  for (int i = tid; i < nq * per_q; i += THREADS) {
    const int qi = i / per_q, r = i - qi * per_q;
    const int w = r / k, j = r - w * k;
    const uint64_t key = lists[((size_t)w * QT + qi) * k + j];
    if (key == SB_KEY_EMPTY) continue;
    int rank = 0;
#pragma unroll
    for (int ww = 0; ww < WARPS; ++ww) {
      const uint64_t* l = lists + ((size_t)ww * QT + qi) * k;
      int lo = 0, hi = k;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (l[mid] < key) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) part[((size_t)(qt0 + qi) * P + blockIdx.x) * k + rank] = key;
  }
}

template <int W>
struct ScanCfg {
  static constexpr int C = (W <= 4) ? 8 : (W == 8 ? 4 : (W == 16 ? 2 : 1));  // codes per lane
  static constexpr int TILE = 32 * C;                                          // codes per warp tile
};

template <int W, int MODE>
__global__ void __launch_bounds__(THREADS, 2)
hamming_scan_kernel(const uint32_t* __restrict__ db, long long U, const uint32_t* __restrict__ qcodes, int Q, int QT,
                    int k, long long idx_base, long long codes_per_chunk, uint64_t* __restrict__ part, int P,
                    int* __restrict__ tau_g, int* __restrict__ hist_g, int q_tma_ok, int one,
                    const int* __restrict__ enable) {
  if (enable != nullptr && *enable == 0) return;       // predicated fallback (sb_hamming_scan_if): nothing to redo
  constexpr int C = ScanCfg<W>::C;
  constexpr int TILE = ScanCfg<W>::TILE;
  constexpr int MAXD = 32 * W;       // largest possible distance
  constexpr int HB = MAXD + 32;      // histogram row pitch (bins 0..MAXD, padded)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [mbarrier 16B][queries QT*W u32][tau QT i32][lists WARPS*QT*k u64]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* sq = reinterpret_cast<uint32_t*>(smem_raw + 16);
  int* stau = reinterpret_cast<int*>(sq + (size_t)QT * W);
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw + 16 + (((size_t)QT * W * 4 + (size_t)QT * 4 + 7) / 8) * 8);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qt0 = blockIdx.y * QT;
  const int nq = min(QT, Q - qt0);

  // -- stage the query tile (TMA bulk copy for the 16-byte-multiple body) --
  const uint32_t qbytes = (uint32_t)nq * W * 4u;
  const uint32_t bulk = q_tma_ok ? (qbytes & ~15u) : 0u;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, bulk);
    if (bulk) tma_bulk_g2s(sq, qcodes + (size_t)qt0 * W, bulk, bar);
  }
  for (uint32_t i = bulk / 4 + tid; i < qbytes / 4; i += THREADS) sq[i] = qcodes[(size_t)qt0 * W + i];
  for (int i = tid; i < QT; i += THREADS) stau[i] = (i < nq) ? min(TAU_INIT, tau_g[qt0 + i]) : TAU_INIT;
  for (int i = tid; i < WARPS * QT * k; i += THREADS) lists[i] = SB_KEY_EMPTY;
  mbar_wait(bar, 0);
  __syncthreads();

  const long long chunk_begin = (long long)blockIdx.x * codes_per_chunk;
  const long long chunk_end = min(U, chunk_begin + codes_per_chunk);
  uint64_t* mylists = lists + (size_t)warp * QT * k;
  volatile int* vtau = stau;

  for (long long tb = chunk_begin + (long long)warp * TILE; tb < chunk_end; tb += (long long)WARPS * TILE) {
    uint32_t code[C][W];
    const bool full = (tb + TILE <= chunk_end);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      long long r = tb + c * 32 + lane;
      if (full || r < chunk_end) {
        load_words<W>(code[c], db + (size_t)r * W);
      } else {
#pragma unroll
        for (int i = 0; i < W; ++i) code[c][i] = 0u;
      }
    }
    // pull in thresholds published by other CTAs (any full list's k-th distance
    // is a valid upper bound for every list of the same query)
    for (int i = lane; i < nq; i += 32) {
      const int g = __ldcg(tau_g + qt0 + i);
      if (g < vtau[i]) atomicMin(&stau[i], g);
    }
    const long long row0 = idx_base + tb + lane;
    unsigned valid = 0xffffffffu;  // bit c: this lane's code c exists
    if (!full) {
      valid = 0u;
#pragma unroll
      for (int c = 0; c < C; ++c) valid |= (tb + c * 32 + lane < chunk_end) ? (1u << c) : 0u;
    }

    // Two queries per trip: one warp vote (and one dependent branch) per 2*C pairs.
    auto scan_queries = [&](auto full_c) {
    constexpr bool FULL = decltype(full_c)::value;              // no per-pair tail masking on whole tiles
    for (int qi = 0; qi < nq; qi += 2) {
      const bool two = (qi + 1 < nq);
      uint32_t qa[W], qb[W];
      load_words_shared<W>(qa, sq + (size_t)qi * W);
      load_words_shared<W>(qb, sq + (size_t)(two ? qi + 1 : qi) * W);
      const int tau_a = vtau[qi];
      const int tau_b = two ? vtau[qi + 1] : -1;
      int da[C], dbb[C];
      bool any = false;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        da[c] = hamming<W, MODE>(code[c], qa, one);
        dbb[c] = hamming<W, MODE>(code[c], qb, one);
        if (!FULL) {
          const bool ok = (valid >> c) & 1u;
          da[c] = ok ? da[c] : D_INVALID;
          dbb[c] = ok ? dbb[c] : D_INVALID;
        }
        any |= (da[c] <= tau_a) | (dbb[c] <= tau_b);
      }
      if (__any_sync(sb::FULL_MASK, any)) {
        // ---- slow path: warp-cooperative insertion (rare after warm-up) ----
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          if (h == 1 && !two) break;
          uint64_t* list = mylists + (size_t)(qi + h) * k;
          const int tau = h ? tau_b : tau_a;
          bool touched = false;
          // Global distance histogram of passing candidates (all CTAs).  If k distinct
          // rows with distance <= D have been seen anywhere, D bounds the k-th distance.
          // Under-counting (skipped while tau is still unset) only loosens the bound.
          int* hrow = hist_g + (size_t)(qt0 + qi + h) * HB;
          const bool hist_on = (tau <= MAXD);
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const int d = h ? dbb[c] : da[c];
            if (hist_on && d <= tau) atomicAdd(hrow + d, 1);
            const uint64_t mykey = ((uint64_t)(uint32_t)d << SB_KEY_ROW_BITS) | (uint64_t)(row0 + c * 32);
            // candidates must beat this warp's own k-th key; re-evaluated after each insert
            unsigned m = __ballot_sync(sb::FULL_MASK, d <= tau && mykey < list[k - 1]);
            while (m) {
              const int src = __ffs(m) - 1;
              const uint64_t key = __shfl_sync(sb::FULL_MASK, mykey, src);
              warp_insert(list, k, key, lane);
              touched = true;
              m &= m - 1;
              m &= __ballot_sync(sb::FULL_MASK, mykey < list[k - 1]);
            }
          }
          int bound = tau;
          if (touched) {
            const uint64_t kth = list[k - 1];
            if (kth != SB_KEY_EMPTY) bound = min(bound, (int)(kth >> SB_KEY_ROW_BITS));
          }
          if (hist_on) {
            // smallest D in the 32-bin window below tau whose cumulative count reaches k
            const int base = max(tau - 31, 0);
            const int bin = base + lane;
            int cum = (bin <= tau) ? __ldcg(hrow + bin) : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int t = __shfl_up_sync(sb::FULL_MASK, cum, o);
              if (lane >= o) cum += t;
            }
            const unsigned ok = __ballot_sync(sb::FULL_MASK, cum >= k);
            if (ok) bound = min(bound, base + __ffs(ok) - 1);
          }
          if (lane == 0 && bound < tau) {
            if (bound < atomicMin(&stau[qi + h], bound)) atomicMin(tau_g + qt0 + qi + h, bound);
          }
        }
      }
    }
    };
    if (full) scan_queries(std::true_type{}); else scan_queries(std::false_type{});

  }

  fold_lists(lists, QT, nq, k, part, P, qt0, tid);
}

// ---- few-queries scan (Q <= FEW_Q): the LinearHashIndex.nn call shape ---------------
// One query does ~20 integer ops per 32-byte code, so this regime is HBM-bound: the
// table is streamed exactly once.  Same chunk / warp-tile decomposition and per-warp
// k-lists as the batched kernel, but every warp keeps TWO tiles of codes in flight
// (register double buffer): the loads of tile i+1 are issued before the XOR/POPC of
// tile i, which is what keeps enough bytes in flight to cover HBM latency.
constexpr int FEW_Q = 4;

template <int W>
__global__ void __launch_bounds__(THREADS, 2)
hamming_scan_few_kernel(const uint32_t* __restrict__ db, long long U, const uint32_t* __restrict__ qcodes, int Q, int k,
                        long long idx_base, long long codes_per_chunk, uint64_t* __restrict__ part, int P,
                        int* __restrict__ tau_g, const int* __restrict__ enable) {
  if (enable != nullptr && *enable == 0) return;
  constexpr int C = ScanCfg<W>::C;
  constexpr int TILE = ScanCfg<W>::TILE;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [queries FEW_Q*W u32][tau FEW_Q i32][lists WARPS*FEW_Q*k u64]
  uint32_t* sq = reinterpret_cast<uint32_t*>(smem_raw);
  int* stau = reinterpret_cast<int*>(sq + FEW_Q * W);
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw + (((size_t)FEW_Q * W * 4 + FEW_Q * 4 + 7) / 8) * 8);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nq = Q;

  for (int i = tid; i < nq * W; i += THREADS) sq[i] = qcodes[i];
  for (int i = tid; i < FEW_Q; i += THREADS) stau[i] = (i < nq) ? min(TAU_INIT, tau_g[i]) : TAU_INIT;
  for (int i = tid; i < WARPS * FEW_Q * k; i += THREADS) lists[i] = SB_KEY_EMPTY;
  __syncthreads();

  const long long chunk_begin = (long long)blockIdx.x * codes_per_chunk;
  const long long chunk_end = min(U, chunk_begin + codes_per_chunk);
  uint64_t* mylists = lists + (size_t)warp * FEW_Q * k;
  volatile int* vtau = stau;

  auto load_tile = [&](uint32_t (&dst)[C][W], long long tb) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const long long r = tb + c * 32 + lane;
      if (r < chunk_end) {
        load_words<W>(dst[c], db + (size_t)r * W);
      } else {
#pragma unroll
        for (int i = 0; i < W; ++i) dst[c][i] = 0u;
      }
    }
  };

  uint32_t cur[C][W], nxt[C][W];
  long long tb = chunk_begin + (long long)warp * TILE;
  if (tb < chunk_end) load_tile(cur, tb);
  while (tb < chunk_end) {
    const long long tn = tb + (long long)WARPS * TILE;
    if (tn < chunk_end) load_tile(nxt, tn);                     // in flight while `cur` is processed
    if (lane < nq) {
      const int g = __ldcg(tau_g + lane);
      if (g < vtau[lane]) atomicMin(&stau[lane], g);
    }
    const long long row0 = idx_base + tb + lane;
    for (int qi = 0; qi < nq; ++qi) {
      uint32_t qw[W];
      load_words_shared<W>(qw, sq + (size_t)qi * W);
      const int tau = vtau[qi];
      int d[C];
      bool any = false;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        d[c] = hamming<W, 1>(cur[c], qw, 1);
        if (tb + c * 32 + lane >= chunk_end) d[c] = D_INVALID;
        any |= (d[c] <= tau);
      }
      if (__any_sync(sb::FULL_MASK, any)) {
        uint64_t* list = mylists + (size_t)qi * k;
        bool touched = false;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const uint64_t mykey = ((uint64_t)(uint32_t)d[c] << SB_KEY_ROW_BITS) | (uint64_t)(row0 + c * 32);
          unsigned m = __ballot_sync(sb::FULL_MASK, d[c] <= tau && mykey < list[k - 1]);
          while (m) {
            const int src = __ffs(m) - 1;
            const uint64_t key = __shfl_sync(sb::FULL_MASK, mykey, src);
            warp_insert(list, k, key, lane);
            touched = true;
            m &= m - 1;
            m &= __ballot_sync(sb::FULL_MASK, mykey < list[k - 1]);
          }
        }
        if (touched) {
          const uint64_t kth = list[k - 1];
          if (lane == 0 && kth != SB_KEY_EMPTY) {
            const int bound = (int)(kth >> SB_KEY_ROW_BITS);
            if (bound < tau && bound < atomicMin(&stau[qi], bound)) atomicMin(tau_g + qi, bound);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int i = 0; i < W; ++i) cur[c][i] = nxt[c][i];
    tb = tn;
  }
  fold_lists(lists, FEW_Q, nq, k, part, P, 0, tid);
}

size_t few_smem_bytes(int W, int k) {
  return (((size_t)FEW_Q * W * 4 + FEW_Q * 4 + 7) / 8) * 8 + (size_t)WARPS * FEW_Q * k * 8;
}

// ---- threshold seeding ----------------------------------------------------------
// One warp per query scans a small prefix sample of the table and publishes the
// sample's k-th smallest distance as the query's initial threshold.  The sample
// is part of the table, so its k-th distance bounds the table's k-th distance.
constexpr int SEED_WARPS = 4;

template <int W>
__global__ void __launch_bounds__(SEED_WARPS * 32)
tau_seed_kernel(const uint32_t* __restrict__ db, int S, const uint32_t* __restrict__ qcodes, int Q, int k,
                int* __restrict__ tau_g, const int* __restrict__ enable) {
  if (enable != nullptr && *enable == 0) return;
  constexpr int MAXD = 32 * W;
  extern __shared__ int seed_hist[];  // [SEED_WARPS][MAXD + 1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * SEED_WARPS + warp;
  int* h = seed_hist + warp * (MAXD + 1);
  for (int i = lane; i <= MAXD; i += 32) h[i] = 0;
  __syncwarp();
  if (q >= Q) return;
  uint32_t qw[W];
#pragma unroll
  for (int i = 0; i < W; ++i) qw[i] = qcodes[(size_t)q * W + i];
  for (int r = lane; r < S; r += 32) {
    uint32_t c[W];
    load_words<W>(c, db + (size_t)r * W);
    atomicAdd(&h[hamming<W, 1>(c, qw, 1)], 1);
  }
  __syncwarp();
  // k-th smallest distance of the sample
  int acc = 0, found = -1;
  for (int base = 0; base <= MAXD && found < 0; base += 32) {
    const int bin = base + lane;
    int cum = (bin <= MAXD) ? h[bin] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(sb::FULL_MASK, cum, o);
      if (lane >= o) cum += t;
    }
    const unsigned ok = __ballot_sync(sb::FULL_MASK, acc + cum >= k);
    if (ok) found = base + __ffs(ok) - 1;
    acc += __shfl_sync(sb::FULL_MASK, cum, 31);
  }
  if (lane == 0 && found >= 0) tau_g[q] = found;
}

// Cheaper seeding for k <= 32: no histogram, no atomics.  Every lane keeps the smallest
// distance among the sample rows it visits; the 32 lane minima belong to 32 distinct
// rows, so their k-th smallest bounds the table's k-th distance.  With S rows per query
// this is as tight as the exact k-th of the sample (both sit near the k/S quantile).
template <int W>
__global__ void __launch_bounds__(SEED_WARPS * 32)
tau_seed_min_kernel(const uint32_t* __restrict__ db, int S, const uint32_t* __restrict__ qcodes, int Q, int k,
                    int* __restrict__ tau_g, const int* __restrict__ enable) {
  if (enable != nullptr && *enable == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * SEED_WARPS + warp;
  if (q >= Q) return;
  uint32_t qw[W];
#pragma unroll
  for (int i = 0; i < W; ++i) qw[i] = qcodes[(size_t)q * W + i];
  int best = D_INVALID;
#pragma unroll 4
  for (int r = lane; r < S; r += 32) {
    uint32_t c[W];
    load_words<W>(c, db + (size_t)r * W);
    best = min(best, hamming<W, 1>(c, qw, 1));
  }
  // rank of this lane's minimum among the 32 (ties by lane): the lane of rank k-1 publishes
  int rank = 0;
#pragma unroll
  for (int o = 0; o < 32; ++o) {
    const int other = __shfl_sync(sb::FULL_MASK, best, o);
    rank += (other < best || (other == best && o < lane)) ? 1 : 0;
  }
  if (rank == k - 1 && best != D_INVALID) tau_g[q] = best;
}

// ---- merge: P sorted k-lists per query -> top-k ------------------------------
// Exact selection by distance histogram: distances are small integers (<= 32*W
// <= 1024), so one pass counts keys per distance, a scan finds the distance d*
// of the k-th key, and a second pass keeps every key below d* plus the
// lowest-row keys at d*.  Survivors (normally ~k) are ordered by counting.
constexpr int MERGE_THREADS = 256;
constexpr int MERGE_CAP = 4096;   // survivor keys staged in shared memory
constexpr int MERGE_BINS = 1032;  // distances 0..1024 (+ padding)

// lists: u64[P][Q][k] when q_major == 0 (all-gathered per-device results), or
// u64[Q][P][k] when q_major == 1 (scan workspace).
__global__ void __launch_bounds__(MERGE_THREADS)
merge_kernel(const uint64_t* __restrict__ lists, int P, int Q, int k, int q_major, uint64_t* __restrict__ out_keys,
             int32_t* __restrict__ out_dist, long long* __restrict__ out_idx, const int* __restrict__ enable) {
  if (enable != nullptr && *enable == 0) return;
  extern __shared__ __align__(8) unsigned char merge_dyn[];
  uint64_t* s_cand = reinterpret_cast<uint64_t*>(merge_dyn);          // [MERGE_CAP]
  uint64_t* s_out = s_cand + MERGE_CAP;                               // [k]
  int* s_hist = reinterpret_cast<int*>(s_out + k);                    // [MERGE_BINS]
  __shared__ int s_count, s_dstar, s_below;
  __shared__ unsigned long long s_cut;  // largest row id taken at distance d*

  const int q = blockIdx.x, tid = threadIdx.x;
  const size_t list_stride = q_major ? (size_t)k : (size_t)Q * k;
  const uint64_t* base = q_major ? lists + (size_t)q * P * k : lists + (size_t)q * k;
  const int total = P * k;
  const uint64_t row_mask = (1ull << SB_KEY_ROW_BITS) - 1;

  for (int i = tid; i < MERGE_BINS; i += MERGE_THREADS) s_hist[i] = 0;
  for (int i = tid; i < k; i += MERGE_THREADS) s_out[i] = SB_KEY_EMPTY;
  if (tid == 0) { s_count = 0; s_cut = row_mask; }
  __syncthreads();

  // 1. histogram of distances
  for (int i = tid; i < total; i += MERGE_THREADS) {
    const int p = i / k, j = i - p * k;
    const uint64_t key = base[p * list_stride + j];
    if (key != SB_KEY_EMPTY) atomicAdd(&s_hist[min((int)(key >> SB_KEY_ROW_BITS), MERGE_BINS - 1)], 1);
  }
  __syncthreads();
  // 2. d* = distance of the k-th key (or "everything" when fewer than k keys exist)
  if (tid == 0) {
    int acc = 0, d = 0;
    for (; d < MERGE_BINS; ++d) {
      if (acc + s_hist[d] >= k) break;
      acc += s_hist[d];
    }
    s_dstar = d;       // MERGE_BINS when fewer than k keys in total
    s_below = acc;     // keys strictly below d*
  }
  __syncthreads();
  const int dstar = s_dstar, below = s_below;
  const int need_at = k - below;                                     // keys wanted at distance d*
  const int have_at = (dstar < MERGE_BINS) ? s_hist[dstar] : 0;

  // 2b. massive ties at d*: bisect the row id so that exactly need_at keys remain.
  if (dstar < MERGE_BINS && below + have_at > MERGE_CAP) {
    unsigned long long lo = 0, hi = row_mask;                        // smallest x with count(row <= x) >= need_at
    while (lo < hi) {
      const unsigned long long mid = lo + ((hi - lo) >> 1);
      __syncthreads();
      if (tid == 0) s_count = 0;
      __syncthreads();
      int c = 0;
      for (int i = tid; i < total; i += MERGE_THREADS) {
        const int p = i / k, j = i - p * k;
        const uint64_t key = base[p * list_stride + j];
        if (key != SB_KEY_EMPTY && (int)(key >> SB_KEY_ROW_BITS) == dstar && (key & row_mask) <= mid) ++c;
      }
      if (c) atomicAdd(&s_count, c);
      __syncthreads();
      if (s_count >= need_at) hi = mid; else lo = mid + 1;
    }
    __syncthreads();
    if (tid == 0) { s_cut = lo; s_count = 0; }
    __syncthreads();
  }
  const unsigned long long cut = s_cut;

  // 3. collect survivors
  for (int i = tid; i < total; i += MERGE_THREADS) {
    const int p = i / k, j = i - p * k;
    const uint64_t key = base[p * list_stride + j];
    if (key == SB_KEY_EMPTY) continue;
    const int d = (int)(key >> SB_KEY_ROW_BITS);
    if (d < dstar || (d == dstar && (key & row_mask) <= cut)) {
      const int slot = atomicAdd(&s_count, 1);
      if (slot < MERGE_CAP) s_cand[slot] = key;
    }
  }
  __syncthreads();
  const int n = min(s_count, MERGE_CAP);
  // 4. order by counting (keys are unique)
  for (int i = tid; i < n; i += MERGE_THREADS) {
    const uint64_t key = s_cand[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (s_cand[j] < key);
    if (rank < k) s_out[rank] = key;
  }
  __syncthreads();
  for (int i = tid; i < k; i += MERGE_THREADS) {
    const uint64_t key = s_out[i];
    const size_t o = (size_t)q * k + i;
    if (out_keys) out_keys[o] = key;
    if (out_dist) out_dist[o] = (key == SB_KEY_EMPTY) ? -1 : (int32_t)(key >> SB_KEY_ROW_BITS);
    if (out_idx) out_idx[o] = (key == SB_KEY_EMPTY) ? -1ll : (long long)(key & row_mask);
  }
}

// One list per query (a single part): nothing to merge, the keys are already the answer -- decode them.
__global__ void __launch_bounds__(256)
decode_keys_kernel(const uint64_t* __restrict__ keys, long long n, uint64_t* __restrict__ out_keys, int32_t* __restrict__ out_dist,
                   long long* __restrict__ out_idx, const int* __restrict__ enable) {
  if (enable != nullptr && *enable == 0) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t key = keys[i];
  const uint64_t row_mask = (1ull << SB_KEY_ROW_BITS) - 1;
  if (out_keys) out_keys[i] = key;
  if (out_dist) out_dist[i] = (key == SB_KEY_EMPTY) ? -1 : (int32_t)(key >> SB_KEY_ROW_BITS);
  if (out_idx) out_idx[i] = (key == SB_KEY_EMPTY) ? -1ll : (long long)(key & row_mask);
}

size_t merge_smem_bytes(int k) { return (size_t)MERGE_CAP * 8 + (size_t)k * 8 + (size_t)MERGE_BINS * 4; }

struct ScanPlan {
  int QT, nqt, chunks, P;
  long long codes_per_chunk;
  size_t smem_bytes, tau_bytes, hist_bytes, workspace_bytes;  // workspace = [tau_g][hist_g][part lists]
};

size_t scan_smem_bytes(int W, int QT, int k) {
  return 16 + (((size_t)QT * W * 4 + (size_t)QT * 4 + 7) / 8) * 8 + (size_t)WARPS * QT * k * 8;
}

int tile_codes(int W) {
  switch (W) {
    case 1: return ScanCfg<1>::TILE;
    case 2: return ScanCfg<2>::TILE;
    case 4: return ScanCfg<4>::TILE;
    case 8: return ScanCfg<8>::TILE;
    case 16: return ScanCfg<16>::TILE;
    default: return ScanCfg<32>::TILE;
  }
}

// Deterministic launch plan shared by the workspace query and the launch.
ScanPlan make_plan(long long U, int W, int Q, int k) {
  ScanPlan p;
  const int sms = sb::sm_count();
  // query tile: as large as the per-CTA list budget allows (2 CTAs / SM)
  const size_t list_budget = 96 * 1024;
  int qt_cap = (int)(list_budget / ((size_t)WARPS * k * 8));
  qt_cap = max(1, min(128, qt_cap));
  p.nqt = (Q + qt_cap - 1) / qt_cap;
  p.QT = (Q + p.nqt - 1) / p.nqt;                       // balance the tiles
  if (qt_cap >= 4) p.QT = min(qt_cap & ~3, ((p.QT + 3) / 4) * 4);  // keep tile offsets 16-byte aligned
  p.nqt = (Q + p.QT - 1) / p.QT;
  // code chunks: up to 4 waves of 2 CTAs/SM when there is enough work (measured best at
  // U = 1.25M, 2.5M and 10M rows x 4096 queries), never less than one CTA pass
  // (WARPS warp tiles) per chunk
  const long long pass = (long long)WARPS * tile_codes(W);
  const long long max_chunks = max(1ll, (U + pass - 1) / pass);
  // Grid = nqt * chunks CTAs.  Keep it at (just under) a whole number of waves of
  // the 2-CTA/SM residency so no SM idles behind a partial last wave; thresholds
  // are global, so extra waves cost almost nothing and shorten the tail.
  const long long slots = (long long)sms * 2;
  const double total_pairs = (double)U * (double)Q;
  int waves = (int)(total_pairs / ((double)slots * 2.0e6));
  waves = max(1, min(4, waves));
  if (const char* e = getenv("SB_SCAN_WAVES")) waves = max(1, atoi(e));   // tuning knob
  long long chunks = (slots * waves) / p.nqt;          // floor: never spill into an extra wave
  if (chunks < 1) chunks = 1;
  chunks = min(chunks, max_chunks);
  // round the chunk length up to whole CTA passes so only the last chunk has a tail
  long long cpc = (U + chunks - 1) / chunks;
  cpc = ((cpc + pass - 1) / pass) * pass;
  if (cpc < pass) cpc = pass;
  chunks = max(1ll, (U + cpc - 1) / cpc);              // rounding up cpc can only lower this
  p.chunks = (int)chunks;
  p.codes_per_chunk = cpc;
  p.P = p.chunks;
  p.smem_bytes = scan_smem_bytes(W, p.QT, k);
  p.tau_bytes = (((size_t)Q * sizeof(int) + 255) / 256) * 256;
  p.hist_bytes = (((size_t)Q * (32 * W + 32) * sizeof(int) + 255) / 256) * 256;
  p.workspace_bytes = p.tau_bytes + p.hist_bytes + (size_t)Q * p.P * k * sizeof(uint64_t);
  return p;
}

template <int W>
int launch_scan(const ScanPlan& p, int mode, const uint32_t* db, long long U, const uint32_t* q, int Q, int k,
                long long idx_base, uint64_t* part, int* tau_g, int* hist_g, const int* enable, cudaStream_t st) {
  if (k <= 32 && U >= 32) {
    // lane-minima seed over a table prefix (the prefix is part of the table, so any k of
    // its rows bound the k-th distance from above)
    long long S = min(U, 8192ll);
    if (const char* e = getenv("SB_SEED_ROWS")) S = min(U, max((long long)atoi(e), 32ll));   // tuning knob
    sb::ProfScope prof("tau_seed_kernel", st);
    tau_seed_min_kernel<W><<<(Q + SEED_WARPS - 1) / SEED_WARPS, SEED_WARPS * 32, 0, st>>>(db, (int)S, q, Q, k, tau_g, enable);
    sb::count_launch();
    int rc = sb::check_launch("tau_seed_min_kernel");
    if (rc) return rc;
  } else {
    // sample = table prefix, large enough to hold k rows several times over
    const long long S = min(U, max(2048ll, 4ll * k));
    if (S >= k) {
      const size_t sh = (size_t)SEED_WARPS * (32 * W + 1) * sizeof(int);
      sb::ProfScope prof("tau_seed_kernel", st);
      tau_seed_kernel<W><<<(Q + SEED_WARPS - 1) / SEED_WARPS, SEED_WARPS * 32, sh, st>>>(db, (int)S, q, Q, k, tau_g, enable);
      sb::count_launch();
      int rc = sb::check_launch("tau_seed_kernel");
      if (rc) return rc;
    }
  }
  if (Q <= FEW_Q && few_smem_bytes(W, k) <= 100 * 1024) {
    const size_t sh = few_smem_bytes(W, k);
    SB_CUDA_TRY(cudaFuncSetAttribute(hamming_scan_few_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    sb::ProfScope prof("hamming_scan_kernel", st);
    hamming_scan_few_kernel<W><<<p.chunks, THREADS, sh, st>>>(db, U, q, Q, k, idx_base, p.codes_per_chunk, part, p.P,
                                                              tau_g, enable);
    sb::count_launch();
    return sb::check_launch("hamming_scan_few_kernel");
  }
  dim3 grid(p.chunks, p.nqt);
  const int q_tma_ok = ((reinterpret_cast<uintptr_t>(q) & 15u) == 0 && ((size_t)p.QT * W * 4) % 16 == 0) ? 1 : 0;
  sb::ProfScope prof("hamming_scan_kernel", st);
#define SB_SCAN_LAUNCH(M)                                                                                          \
  do {                                                                                                            \
    SB_CUDA_TRY(cudaFuncSetAttribute(hamming_scan_kernel<W, M>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                     (int)p.smem_bytes));                                                        \
    hamming_scan_kernel<W, M><<<grid, THREADS, p.smem_bytes, st>>>(db, U, q, Q, p.QT, k, idx_base,                \
                                                                     p.codes_per_chunk, part, p.P, tau_g, hist_g, \
                                                                     q_tma_ok, /*one=*/1, enable);                \
  } while (0)
  if (mode == 0) SB_SCAN_LAUNCH(0);
  else if (mode == 1) SB_SCAN_LAUNCH(1);
  else SB_SCAN_LAUNCH(2);
#undef SB_SCAN_LAUNCH
  sb::count_launch();
  return sb::check_launch("hamming_scan_kernel");
}

int launch_merge(const uint64_t* lists, int P, int Q, int k, int q_major, uint64_t* out_keys, int32_t* out_dist,
                 int64_t* out_idx, cudaStream_t st, const int* enable = nullptr, bool sorted_lists = false) {
  if (P == 1 && sorted_lists) {                               // (the scan's own per-chunk lists are unordered)
    sb::ProfScope prof("decode_keys_kernel", st);
    const long long n = (long long)Q * k;
    decode_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(lists, n, out_keys, out_dist,
                                                                    reinterpret_cast<long long*>(out_idx), enable);
    sb::count_launch();
    return sb::check_launch("decode_keys_kernel");
  }
  const size_t dyn = merge_smem_bytes(k);
  SB_CUDA_TRY(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  sb::ProfScope prof("merge_kernel", st);
  merge_kernel<<<Q, MERGE_THREADS, dyn, st>>>(lists, P, Q, k, q_major, out_keys, out_dist,
                                              reinterpret_cast<long long*>(out_idx), enable);
  sb::count_launch();
  return sb::check_launch("merge_kernel");
}

int scan_impl(const uint32_t* db, int64_t U, int32_t W, const uint32_t* q, int32_t Q, int32_t k, int64_t idx_base,
              uint64_t* keys_out, int32_t* out_dist, int64_t* out_idx, void* workspace, size_t workspace_bytes,
              int variant, void* stream, const int* enable = nullptr) {
  SB_REQUIRE(W == 1 || W == 2 || W == 4 || W == 8 || W == 16 || W == 32, "sb_hamming_scan: W=%d not in {1,2,4,8,16,32}", W);
  SB_REQUIRE(Q >= 1 && k >= 1 && U >= 0, "sb_hamming_scan: need Q>=1, k>=1, U>=0 (Q=%d k=%d U=%lld)", Q, k, (long long)U);
  SB_REQUIRE(k <= 2048, "sb_hamming_scan: k=%d exceeds the supported maximum of 2048", k);
  SB_REQUIRE(idx_base >= 0 && idx_base + U < (1ll << SB_KEY_ROW_BITS), "sb_hamming_scan: row ids exceed 2^40");
  SB_REQUIRE(q != nullptr && (U == 0 || db != nullptr), "sb_hamming_scan: null input pointer");
  SB_REQUIRE((reinterpret_cast<uintptr_t>(db) & 15u) == 0, "sb_hamming_scan: db must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ScanPlan p = make_plan(U, W, Q, k);
  if (workspace_bytes < p.workspace_bytes || (p.workspace_bytes && workspace == nullptr)) {
    sb::set_error("sb_hamming_scan: workspace too small (%zu < %zu bytes)", workspace_bytes, p.workspace_bytes);
    return SB_ERR_WORKSPACE;
  }
  if (p.smem_bytes > 227 * 1024) {
    sb::set_error("sb_hamming_scan: k=%d needs %zu bytes of shared memory per CTA", k, p.smem_bytes);
    return SB_ERR_UNSUPPORTED;
  }
  int* tau_g = reinterpret_cast<int*>(workspace);
  int* hist_g = reinterpret_cast<int*>(static_cast<unsigned char*>(workspace) + p.tau_bytes);
  uint64_t* part = reinterpret_cast<uint64_t*>(static_cast<unsigned char*>(workspace) + p.tau_bytes + p.hist_bytes);
  SB_CUDA_TRY(cudaMemsetAsync(tau_g, 0x7f, p.tau_bytes, st));  // 0x7f7f7f7f: "no bound yet"
  SB_CUDA_TRY(cudaMemsetAsync(hist_g, 0, p.hist_bytes, st));
  // variant 0 (default) = 4-POPC adder tree, 1 = plain POPC per word, 2 = 5-POPC adder tree
  const int mode = (variant == 1) ? 0 : (variant == 2 ? 1 : 2);
  int rc;
  switch (W) {
    case 1: rc = launch_scan<1>(p, mode, db, U, q, Q, k, idx_base, part, tau_g, hist_g, enable, st); break;
    case 2: rc = launch_scan<2>(p, mode, db, U, q, Q, k, idx_base, part, tau_g, hist_g, enable, st); break;
    case 4: rc = launch_scan<4>(p, mode, db, U, q, Q, k, idx_base, part, tau_g, hist_g, enable, st); break;
    case 8: rc = launch_scan<8>(p, mode, db, U, q, Q, k, idx_base, part, tau_g, hist_g, enable, st); break;
    case 16: rc = launch_scan<16>(p, mode, db, U, q, Q, k, idx_base, part, tau_g, hist_g, enable, st); break;
    default: rc = launch_scan<32>(p, mode, db, U, q, Q, k, idx_base, part, tau_g, hist_g, enable, st); break;
  }
  if (rc) return rc;
  return launch_merge(part, p.P, Q, k, /*q_major=*/1, keys_out, out_dist, out_idx, st, enable);
}

}  // namespace

extern "C" {

size_t sb_hamming_scan_workspace_bytes(int64_t U, int32_t W, int32_t Q, int32_t k) {
  if (!(W == 1 || W == 2 || W == 4 || W == 8 || W == 16 || W == 32) || Q < 1 || k < 1 || U < 0) return 0;
  return make_plan(U, W, Q, k).workspace_bytes;
}

int sb_hamming_scan(const uint32_t* db, int64_t U, int32_t W, const uint32_t* q, int32_t Q, int32_t k,
                    int64_t idx_base, uint64_t* keys_out, void* workspace, size_t workspace_bytes, void* stream) {
  SB_REQUIRE(keys_out != nullptr, "sb_hamming_scan: keys_out is NULL");
  return scan_impl(db, U, W, q, Q, k, idx_base, keys_out, nullptr, nullptr, workspace, workspace_bytes, 0, stream);
}

/* Predicated form: every kernel of the scan returns at once unless *enable (device int) != 0, in which
 * case keys_out is overwritten with the exact result.  Launched right after sb_hamming_scan_tc with its
 * overflow flag, it makes the tensor-core path exact WITHOUT a host round trip for the flag. */
int sb_hamming_scan_if(const uint32_t* db, int64_t U, int32_t W, const uint32_t* q, int32_t Q, int32_t k,
                       int64_t idx_base, uint64_t* keys_out, const int32_t* enable, void* workspace,
                       size_t workspace_bytes, void* stream) {
  SB_REQUIRE(keys_out != nullptr && enable != nullptr, "sb_hamming_scan_if: NULL pointer");
  return scan_impl(db, U, W, q, Q, k, idx_base, keys_out, nullptr, nullptr, workspace, workspace_bytes, 0, stream, enable);
}

int sb_hamming_scan_variant(const uint32_t* db, int64_t U, int32_t W, const uint32_t* q, int32_t Q, int32_t k,
                            int64_t idx_base, uint64_t* keys_out, void* workspace, size_t workspace_bytes,
                            int32_t variant, void* stream) {
  SB_REQUIRE(keys_out != nullptr, "sb_hamming_scan_variant: keys_out is NULL");
  return scan_impl(db, U, W, q, Q, k, idx_base, keys_out, nullptr, nullptr, workspace, workspace_bytes, variant,
                   stream);
}

int sb_hamming_topk(const uint32_t* db, int64_t U, int32_t W, const uint32_t* q, int32_t Q, int32_t k,
                    int64_t idx_base, int32_t* out_dist, int64_t* out_idx, void* workspace, size_t workspace_bytes,
                    void* stream) {
  SB_REQUIRE(out_dist != nullptr && out_idx != nullptr, "sb_hamming_topk: NULL output");
  return scan_impl(db, U, W, q, Q, k, idx_base, nullptr, out_dist, out_idx, workspace, workspace_bytes, 0, stream);
}

int sb_topk_merge(const uint64_t* keys_in, int32_t parts, int32_t Q, int32_t k, uint64_t* out_keys,
                  int32_t* out_dist, int64_t* out_idx, void* stream) {
  SB_REQUIRE(keys_in != nullptr && parts >= 1 && Q >= 1 && k >= 1, "sb_topk_merge: bad arguments");
  return launch_merge(keys_in, parts, Q, k, /*q_major=*/0, out_keys, out_dist, out_idx,
                      reinterpret_cast<cudaStream_t>(stream), nullptr, /*sorted_lists=*/true);
}

}  // extern "C"
