// Index build: sorted-unique code table + code -> rows CSR from per-row packed codes, in ONE device
// pipeline (no host round trip, no library sort).  Replaces the Python set / dict construction of
//   LinearHashIndex._build_index / _update_index / _remove_from_index   smqtk_indexing/impls/hash_index/linear.py:148-204
//   LSHNearestNeighborIndex._build_index hash -> uuids loop              smqtk_indexing/impls/nn_index/lsh.py:316-329
// (a set of Python ints and {int: set(uuid)}), whose array form is: table uint32[U][W] ascending by
// integer value, row -> table row, and the CSR code -> rows with the rows of a code in ascending order.
//
// Pipeline (all on `stream`):
//   1. digit_hist_kernel      one pass over the codes: histogram of EVERY 8-bit digit (4W x 256 bins).
//   2. digit_plan_kernel      exclusive scans -> per-pass digit bases; digits on which all rows agree are
//                             marked "skip" (zero-extended codes, shared prefixes); ping-pong plan.
//   3. radix_pass_kernel x 4W stable LSD radix sort of a uint32 row permutation, least significant
//                             digit first.  One kernel per pass ("onesweep"): a tile of 8192 rows is ranked
//                             with warp match_any (stable: warp chunk, round, lane = position order), the
//                             tile's per-digit offsets come from a decoupled look-back over the
//                             preceding tiles (tiles take tickets, so a tile only ever waits on tiles that
//                             already run), then the permutation is scattered.  The 32-byte rows never
//                             move: a pass gathers one byte per row through the permutation.
//   4. boundary_kernel        flags "differs from its predecessor" (full-row compare) + per-block counts
//   5. block_scan_kernel      exclusive scan of the block counts (one CTA), U
//   6. emit_kernel            table rows, row -> code, csr_off, csr_rows
//   7. max_count_kernel       most rows on one code (sizes the fixed-pitch candidate segments)
// Bound: HBM, ~ (4 + 32 + 4) bytes per row and pass (permutation in, one 32-byte sector per gathered
// byte, permutation out) -- DESIGN.md section 3.6.
#include "common.cuh"

namespace {

constexpr int RADIX = 256;
constexpr int RS_THREADS = 512;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;          // 8192 rows per tile
constexpr uint32_t LB_AGG = 1u << 30, LB_PREFIX = 2u << 30, LB_MASK = (1u << 30) - 1u;
constexpr int SRC_SKIP = -1, SRC_A = 0, SRC_B = 1, SRC_INIT = 2;
constexpr int MAX_PASSES = 128;                          // W <= 32

struct SortCtrl {
  int src[MAX_PASSES];       // where pass p reads its permutation from (SRC_*)
  int ord[MAX_PASSES];       // ordinal among the non-skipped passes
  int tickets[MAX_PASSES];
  int final_src;             // where the sorted permutation ended up
};

__device__ __forceinline__ uint32_t load_perm(int src, const uint32_t* a, const uint32_t* b, const long long* rows_in, long long i) {
  if (src == SRC_A) return a[i];
  if (src == SRC_B) return b[i];
  return rows_in ? (uint32_t)rows_in[i] : (uint32_t)i;
}

// ---- 1. all digit histograms in one pass --------------------------------------------------
// One thread per (row, word); shared histograms for `wpb` words per block pass (<= 8: 32 KB).
__global__ void __launch_bounds__(256)
digit_hist_kernel(const uint32_t* __restrict__ codes, const long long* __restrict__ rows_in, long long n, int W,
                  uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t s_h[];                      // [W * 4][256]
  const int bins = W * 4 * RADIX;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) s_h[i] = 0u;
  __syncthreads();
  const long long total = n * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / W;
    const int w = (int)(i - r * W);
    const long long row = rows_in ? rows_in[r] : r;
    const uint32_t v = __ldg(codes + row * W + w);
    // pass index of byte b of word w: least significant digit first
    const int p0 = (W - 1 - w) * 4;
    atomicAdd(&s_h[(p0 + 0) * RADIX + (v & 255u)], 1u);
    atomicAdd(&s_h[(p0 + 1) * RADIX + ((v >> 8) & 255u)], 1u);
    atomicAdd(&s_h[(p0 + 2) * RADIX + ((v >> 16) & 255u)], 1u);
    atomicAdd(&s_h[(p0 + 3) * RADIX + (v >> 24)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x)
    if (s_h[i]) atomicAdd(hist + i, s_h[i]);
}

// ---- 2. per-pass digit bases, skip flags, ping-pong plan -----------------------------------
__global__ void __launch_bounds__(RADIX)
digit_plan_kernel(uint32_t* __restrict__ hist, long long n, int passes, SortCtrl* __restrict__ ctrl) {
  __shared__ uint32_t s_scan[RADIX];
  __shared__ int s_skip[MAX_PASSES];
  const int d = threadIdx.x;
  for (int p = 0; p < passes; ++p) {
    const uint32_t c = hist[p * RADIX + d];
    if (d == 0) s_skip[p] = 0;
    __syncthreads();
    if ((long long)c == n) s_skip[p] = 1;               // every row has this digit: the pass is the identity
    s_scan[d] = c;
    __syncthreads();
    for (int o = 1; o < RADIX; o <<= 1) {                 // Hillis-Steele inclusive scan
      const uint32_t t = (d >= o) ? s_scan[d - o] : 0u;
      __syncthreads();
      s_scan[d] += t;
      __syncthreads();
    }
    hist[p * RADIX + d] = s_scan[d] - c;                  // exclusive: first output slot of digit d in pass p
    __syncthreads();
  }
  if (d == 0) {
    int cur = SRC_INIT, k = 0;
    for (int p = 0; p < passes; ++p) {
      ctrl->tickets[p] = 0;
      if (s_skip[p]) { ctrl->src[p] = SRC_SKIP; ctrl->ord[p] = k; continue; }
      ctrl->src[p] = cur;
      ctrl->ord[p] = k++;
      cur = (cur == SRC_A) ? SRC_B : SRC_A;               // INIT -> A, A -> B, B -> A
    }
    ctrl->final_src = cur;
  }
}

// ---- 3. one stable radix pass over the permutation -----------------------------------------
__global__ void __launch_bounds__(RS_THREADS)
radix_pass_kernel(const uint32_t* __restrict__ codes, int W, int pass, long long n, const long long* __restrict__ rows_in,
                  uint32_t* __restrict__ perm_a, uint32_t* __restrict__ perm_b, const uint32_t* __restrict__ digit_base,
                  SortCtrl* __restrict__ ctrl, uint32_t* __restrict__ lb_state /* [2][tiles][256] */, int tiles) {
  const int src = ctrl->src[pass];
  if (src == SRC_SKIP) return;
  __shared__ uint32_t s_cnt[RS_WARPS][RADIX];
  __shared__ uint32_t s_base[RADIX];
  __shared__ int s_tile;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_tile = atomicAdd(&ctrl->tickets[pass], 1);
  for (int i = tid; i < RS_WARPS * RADIX; i += RS_THREADS) (&s_cnt[0][0])[i] = 0u;
  __syncthreads();
  const int tile = s_tile;
  const int ord = ctrl->ord[pass];
  uint32_t* state = lb_state + (size_t)(ord & 1) * tiles * RADIX;
  uint32_t* state_next = lb_state + (size_t)((ord + 1) & 1) * tiles * RADIX;
  uint32_t* out = (src == SRC_A) ? perm_b : perm_a;        // INIT and B write A
  const int word = W - 1 - (pass >> 2), shift = (pass & 3) * 8;

  uint32_t p[RS_ITEMS];
  unsigned short rank[RS_ITEMS];
  unsigned char dig[RS_ITEMS];
  const long long base = (long long)tile * RS_TILE + (long long)warp * (RS_ITEMS * 32);
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const long long i = base + j * 32 + lane;
    const bool valid = i < n;
    uint32_t d = 0u;
    if (valid) {
      p[j] = load_perm(src, perm_a, perm_b, rows_in, i);
      d = (__ldg(codes + (size_t)p[j] * W + word) >> shift) & 255u;
    }
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const unsigned peers = __match_any_sync(0xffffffffu, d) & vmask;
    const int leader = valid ? (__ffs(peers) - 1) : lane;
    uint32_t old = 0u;
    if (valid && lane == leader) {
      old = s_cnt[warp][d];
      s_cnt[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[j] = (unsigned short)(old + __popc(peers & lt));
    dig[j] = (unsigned char)d;
    __syncwarp();
  }
  __syncthreads();
  if (tid < RADIX) {
    // exclusive offsets of this digit over the tile's warps, tile total, then the look-back
    uint32_t run = 0u;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const uint32_t c = s_cnt[w][tid];
      s_cnt[w][tid] = run;
      run += c;
    }
    volatile uint32_t* vs = state;
    uint32_t excl = 0u;
    if (tile == 0) {
      vs[tid] = run | LB_PREFIX;
    } else {
      vs[(size_t)tile * RADIX + tid] = run | LB_AGG;
      for (int look = tile - 1; look >= 0; --look) {
        uint32_t v;
        do { v = vs[(size_t)look * RADIX + tid]; } while ((v >> 30) == 0u);
        excl += v & LB_MASK;
        if (v & LB_PREFIX) break;
      }
      vs[(size_t)tile * RADIX + tid] = (excl + run) | LB_PREFIX;
    }
    state_next[(size_t)tile * RADIX + tid] = 0u;           // the other buffer is clean for the next active pass
    s_base[tid] = digit_base[pass * RADIX + tid] + excl;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const long long i = base + j * 32 + lane;
    if (i < n) out[s_base[dig[j]] + s_cnt[warp][dig[j]] + rank[j]] = p[j];
  }
}

// ---- 4. boundaries ---------------------------------------------------------------------------
constexpr int BD_THREADS = 256;
constexpr int BD_ITEMS = 4;
constexpr int BD_TILE = BD_THREADS * BD_ITEMS;

__device__ __forceinline__ bool rows_differ(const uint32_t* __restrict__ codes, int W, uint32_t a, uint32_t b) {
  const uint32_t* x = codes + (size_t)a * W;
  const uint32_t* y = codes + (size_t)b * W;
  bool diff = false;
  for (int w = 0; w < W; ++w) diff |= (__ldg(x + w) != __ldg(y + w));
  return diff;
}

__global__ void __launch_bounds__(BD_THREADS)
boundary_kernel(const uint32_t* __restrict__ codes, int W, long long n, const long long* __restrict__ rows_in,
                const uint32_t* __restrict__ perm_a, const uint32_t* __restrict__ perm_b, const SortCtrl* __restrict__ ctrl,
                unsigned char* __restrict__ flags, uint32_t* __restrict__ block_counts) {
  const int src = ctrl->final_src;
  __shared__ int s_total;
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  int mine = 0;
  const long long i0 = (long long)blockIdx.x * BD_TILE + (long long)threadIdx.x * BD_ITEMS;
  uint32_t prev = 0u;
  if (i0 > 0 && i0 < n) prev = load_perm(src, perm_a, perm_b, rows_in, i0 - 1);
#pragma unroll
  for (int j = 0; j < BD_ITEMS; ++j) {
    const long long i = i0 + j;
    if (i < n) {
      const uint32_t cur = load_perm(src, perm_a, perm_b, rows_in, i);
      const bool is_new = (i == 0) || rows_differ(codes, W, prev, cur);
      flags[i] = is_new ? 1 : 0;
      mine += is_new ? 1 : 0;
      prev = cur;
    }
  }
  for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_total, mine);
  __syncthreads();
  if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)s_total;
}

// ---- 5. exclusive scan of the block counts (one CTA), U -------------------------------------
__global__ void __launch_bounds__(1024)
block_scan_kernel(uint32_t* __restrict__ block_counts, int nblocks, long long n, long long* __restrict__ stats,
                  long long* __restrict__ csr_off) {
  __shared__ uint32_t s_part[1024];
  const int tid = threadIdx.x;
  const int per = (nblocks + 1023) / 1024;
  const int lo = min(nblocks, tid * per), hi = min(nblocks, lo + per);
  uint32_t sum = 0u;
  for (int i = lo; i < hi; ++i) sum += block_counts[i];
  s_part[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const uint32_t t = (tid >= o) ? s_part[tid - o] : 0u;
    __syncthreads();
    s_part[tid] += t;
    __syncthreads();
  }
  uint32_t run = s_part[tid] - sum;
  for (int i = lo; i < hi; ++i) {
    const uint32_t c = block_counts[i];
    block_counts[i] = run;
    run += c;
  }
  if (tid == 1023) {
    stats[0] = (long long)s_part[1023];                    // U
    stats[1] = 0;                                          // max rows per code (max_count_kernel)
    csr_off[s_part[1023]] = n;
  }
}

// ---- 6. emit table / row -> code / CSR -------------------------------------------------------
__global__ void __launch_bounds__(BD_THREADS)
emit_kernel(const uint32_t* __restrict__ codes, int W, long long n, const long long* __restrict__ rows_in,
            const uint32_t* __restrict__ perm_a, const uint32_t* __restrict__ perm_b, const SortCtrl* __restrict__ ctrl,
            const unsigned char* __restrict__ flags, const uint32_t* __restrict__ block_base, uint32_t* __restrict__ table,
            long long* __restrict__ row_code, long long* __restrict__ csr_off, long long* __restrict__ csr_rows) {
  const int src = ctrl->final_src;
  __shared__ uint32_t s_warp[BD_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long i0 = (long long)blockIdx.x * BD_TILE + (long long)tid * BD_ITEMS;
  int f[BD_ITEMS];
  int mine = 0;
#pragma unroll
  for (int j = 0; j < BD_ITEMS; ++j) {
    f[j] = (i0 + j < n) ? (int)flags[i0 + j] : 0;
    mine += f[j];
  }
  int inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = (uint32_t)inc;
  __syncthreads();
  uint32_t woff = 0u;
  for (int w = 0; w < warp; ++w) woff += s_warp[w];
  // number of "new" flags strictly before this thread's first item, over the whole array
  long long code = (long long)block_base[blockIdx.x] + woff + (inc - mine);
#pragma unroll
  for (int j = 0; j < BD_ITEMS; ++j) {
    const long long i = i0 + j;
    if (i >= n) break;
    code += f[j];                                          // inclusive count -> code index = count - 1
    const uint32_t p = load_perm(src, perm_a, perm_b, rows_in, i);
    const long long c = code - 1;
    row_code[p] = c;
    csr_rows[i] = (long long)p;
    if (f[j]) {
      csr_off[c] = i;
      for (int w = 0; w < W; ++w) table[(size_t)c * W + w] = __ldg(codes + (size_t)p * W + w);
    }
  }
}

// ---- 7. most rows on one code ----------------------------------------------------------------
__global__ void __launch_bounds__(256)
max_count_kernel(const long long* __restrict__ csr_off, long long* __restrict__ stats) {
  const long long U = stats[0];
  long long best = 0;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < U; c += (long long)gridDim.x * blockDim.x)
    best = max(best, csr_off[c + 1] - csr_off[c]);
  for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0 && best) atomicMax(reinterpret_cast<unsigned long long*>(stats + 1), (unsigned long long)best);
}

struct UcPlan {
  int passes, tiles, bd_blocks;
  size_t off_hist, off_ctrl, off_state, off_perm_a, off_perm_b, off_flags, off_counts, total;
};

size_t al256(size_t x) { return (x + 255) / 256 * 256; }

UcPlan uc_plan(int64_t n, int32_t W) {
  UcPlan p;
  p.passes = W * 4;
  p.tiles = (int)((n + RS_TILE - 1) / RS_TILE);
  if (p.tiles < 1) p.tiles = 1;
  p.bd_blocks = (int)((n + BD_TILE - 1) / BD_TILE);
  if (p.bd_blocks < 1) p.bd_blocks = 1;
  size_t o = 0;
  p.off_hist = o;   o += al256((size_t)p.passes * RADIX * sizeof(uint32_t));
  p.off_ctrl = o;   o += al256(sizeof(SortCtrl));
  p.off_state = o;  o += al256((size_t)2 * p.tiles * RADIX * sizeof(uint32_t));
  p.off_perm_a = o; o += al256((size_t)n * sizeof(uint32_t));
  p.off_perm_b = o; o += al256((size_t)n * sizeof(uint32_t));
  p.off_flags = o;  o += al256((size_t)n);
  p.off_counts = o; o += al256((size_t)p.bd_blocks * sizeof(uint32_t));
  p.total = o;
  return p;
}

bool uc_supported(int64_t n, int64_t n_rows, int32_t W) {
  return n >= 1 && n <= (1ll << 30) - 1 && n_rows >= n && n_rows <= 0xffffffffll &&
         (W == 1 || W == 2 || W == 4 || W == 8 || W == 16 || W == 32);
}

// hist + plan + 4W radix passes: leaves the sorted permutation in perm_a / perm_b / the initial order (ctrl->final_src)
int sort_rows_stage(const uint32_t* codes, int32_t W, const long long* rin, int64_t n, const UcPlan& p, unsigned char* ws,
                    cudaStream_t st) {
  uint32_t* hist = reinterpret_cast<uint32_t*>(ws + p.off_hist);
  SortCtrl* ctrl = reinterpret_cast<SortCtrl*>(ws + p.off_ctrl);
  uint32_t* state = reinterpret_cast<uint32_t*>(ws + p.off_state);
  uint32_t* perm_a = reinterpret_cast<uint32_t*>(ws + p.off_perm_a);
  uint32_t* perm_b = reinterpret_cast<uint32_t*>(ws + p.off_perm_b);
  // hist, ctrl and both look-back buffers are contiguous at the start of the workspace
  SB_CUDA_TRY(cudaMemsetAsync(ws, 0, p.off_perm_a, st));
  {
    sb::ProfScope prof("digit_hist_kernel", st);
    const size_t smem = (size_t)p.passes * RADIX * sizeof(uint32_t);
    if (smem > 48 * 1024)
      SB_CUDA_TRY(cudaFuncSetAttribute(digit_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = (n * W + 255) / 256 / 8;
    const long long cap = (long long)sb::sm_count() * (smem > 64 * 1024 ? 1 : 4);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    digit_hist_kernel<<<(unsigned)blocks, 256, smem, st>>>(codes, rin, n, W, hist);
    sb::count_launch();
    if (int rc = sb::check_launch("digit_hist_kernel")) return rc;
  }
  digit_plan_kernel<<<1, RADIX, 0, st>>>(hist, n, p.passes, ctrl);
  sb::count_launch();
  if (int rc = sb::check_launch("digit_plan_kernel")) return rc;
  {
    sb::ProfScope prof("radix_pass_kernel", st);            // the 4W passes as one profile record
    for (int pass = 0; pass < p.passes; ++pass) {
      radix_pass_kernel<<<p.tiles, RS_THREADS, 0, st>>>(codes, W, pass, n, rin, perm_a, perm_b, hist, ctrl, state, p.tiles);
      sb::count_launch();
    }
    if (int rc = sb::check_launch("radix_pass_kernel")) return rc;
  }
  return SB_OK;
}

// ---- sorted selection (large candidate lists) and exhaustive Hamming top-k (large k) ----------
__device__ __forceinline__ unsigned long long orderable(double d) {
  // ascending unsigned order == ascending double order; every NaN sorts last
  if (d != d) return ~0ull;
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// key rows u32[M][4] = {segment, dist hi, dist lo, tie}: tie = candidate row (tie_by_row) or 0 (the sort is
// stable: equal keys keep position order).  Entries past cand_cnt (fixed-pitch padding) sort last.
__global__ void __launch_bounds__(256)
select_keys_kernel(const double* __restrict__ dist, const long long* __restrict__ cand_off, const long long* __restrict__ cand_cnt,
                   const long long* __restrict__ cand_idx, int Q, long long M, int tie_by_row, uint32_t* __restrict__ keys) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  int lo = 0, hi = Q;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cand_off[mid] <= j) lo = mid; else hi = mid;
  }
  const bool real = cand_cnt == nullptr || (j - cand_off[lo]) < cand_cnt[lo];
  const unsigned long long k = real ? orderable(dist[j]) : ~0ull;
  uint4 row;
  row.x = (uint32_t)lo;
  row.y = (uint32_t)(k >> 32);
  row.z = (uint32_t)k;
  row.w = (tie_by_row && real) ? (uint32_t)cand_idx[j] : (real ? 0u : 0xffffffffu);
  reinterpret_cast<uint4*>(keys)[j] = row;
}

__global__ void __launch_bounds__(256)
select_emit_kernel(const double* __restrict__ dist, const long long* __restrict__ cand_off, const long long* __restrict__ cand_cnt,
                   const long long* __restrict__ cand_idx, int Q, int n, const uint32_t* __restrict__ perm_a,
                   const uint32_t* __restrict__ perm_b, const SortCtrl* __restrict__ ctrl, long long* __restrict__ out_pos,
                   double* __restrict__ out_dist) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)Q * n) return;
  const int q = (int)(t / n), i = (int)(t % n);
  const long long beg = cand_off[q], m = cand_cnt ? min(cand_cnt[q], cand_off[q + 1] - beg) : cand_off[q + 1] - beg;
  if (i < m) {
    const long long j = (long long)load_perm(ctrl->final_src, perm_a, perm_b, nullptr, beg + i);
    out_pos[t] = cand_idx ? cand_idx[j] : j;
    out_dist[t] = dist[j];
  } else {
    out_pos[t] = -1;
    out_dist[t] = nan("");
  }
}

// key rows u32[Qc * U][4] = {query, (d << 40 | row) hi, lo, 0} for EVERY (query, row) pair
template <int W>
__global__ void __launch_bounds__(256)
ham_all_keys_kernel(const uint32_t* __restrict__ db, long long U, const uint32_t* __restrict__ q, int Qc, long long idx_base,
                    uint32_t* __restrict__ keys) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= U) return;
  uint32_t x[W];
#pragma unroll
  for (int w = 0; w < W; ++w) x[w] = __ldg(db + r * W + w);
  for (int qi = 0; qi < Qc; ++qi) {
    int d = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) d += __popc(x[w] ^ __ldg(q + (long long)qi * W + w));
    const unsigned long long k = ((unsigned long long)(unsigned)d << 40) | (unsigned long long)(idx_base + r);
    reinterpret_cast<uint4*>(keys)[(long long)qi * U + r] = make_uint4((uint32_t)qi, (uint32_t)(k >> 32), (uint32_t)k, 0u);
  }
}

__global__ void __launch_bounds__(256)
ham_all_emit_kernel(const uint32_t* __restrict__ keys, long long U, int Qc, int k, const uint32_t* __restrict__ perm_a,
                    const uint32_t* __restrict__ perm_b, const SortCtrl* __restrict__ ctrl, unsigned long long* __restrict__ keys_out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)Qc * k) return;
  const int qi = (int)(t / k), i = (int)(t % k);
  if (i < U) {
    const long long j = (long long)load_perm(ctrl->final_src, perm_a, perm_b, nullptr, (long long)qi * U + i);
    const uint4 row = reinterpret_cast<const uint4*>(keys)[j];
    keys_out[t] = ((unsigned long long)row.y << 32) | row.z;
  } else {
    keys_out[t] = ~0ull;
  }
}

}  // namespace

extern "C" {

size_t sb_unique_codes_workspace_bytes(int64_t n, int32_t W) {
  if (!uc_supported(n, n, W)) return 0;
  return uc_plan(n, W).total;
}

int sb_unique_codes(const uint32_t* codes, int64_t n_rows, int32_t W, const int64_t* rows_in, int64_t n,
                    uint32_t* table_out, int64_t* row_code_out, int64_t* csr_off_out, int64_t* csr_rows_out,
                    int64_t* stats_out, void* workspace, size_t workspace_bytes, void* stream) {
  SB_REQUIRE(codes && table_out && row_code_out && csr_off_out && csr_rows_out && stats_out, "sb_unique_codes: NULL pointer");
  if (!uc_supported(n, n_rows, W)) {
    sb::set_error("sb_unique_codes: needs 1 <= n <= n_rows, n < 2^30, n_rows < 2^32, W in {1,2,4,8,16,32} (n=%lld n_rows=%lld W=%d)",
                  (long long)n, (long long)n_rows, W);
    return SB_ERR_UNSUPPORTED;
  }
  SB_REQUIRE(rows_in != nullptr || n == n_rows, "sb_unique_codes: n != n_rows needs rows_in");
  const UcPlan p = uc_plan(n, W);
  if (workspace == nullptr || workspace_bytes < p.total) {
    sb::set_error("sb_unique_codes: workspace too small (%zu < %zu bytes)", workspace_bytes, p.total);
    return SB_ERR_WORKSPACE;
  }
  SB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "sb_unique_codes: workspace must be 256-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  SortCtrl* ctrl = reinterpret_cast<SortCtrl*>(ws + p.off_ctrl);
  uint32_t* perm_a = reinterpret_cast<uint32_t*>(ws + p.off_perm_a);
  uint32_t* perm_b = reinterpret_cast<uint32_t*>(ws + p.off_perm_b);
  unsigned char* flags = ws + p.off_flags;
  uint32_t* counts = reinterpret_cast<uint32_t*>(ws + p.off_counts);
  const long long* rin = reinterpret_cast<const long long*>(rows_in);

  if (rows_in != nullptr)                                   // rows outside rows_in (tombstones) map to no code
    SB_CUDA_TRY(cudaMemsetAsync(row_code_out, 0xff, (size_t)n_rows * sizeof(int64_t), st));
  if (int rc = sort_rows_stage(codes, W, rin, n, p, ws, st)) return rc;
  {
    sb::ProfScope prof("boundary_kernel", st);
    boundary_kernel<<<p.bd_blocks, BD_THREADS, 0, st>>>(codes, W, n, rin, perm_a, perm_b, ctrl, flags, counts);
    sb::count_launch();
    if (int rc = sb::check_launch("boundary_kernel")) return rc;
  }
  block_scan_kernel<<<1, 1024, 0, st>>>(counts, p.bd_blocks, n, reinterpret_cast<long long*>(stats_out),
                                        reinterpret_cast<long long*>(csr_off_out));
  sb::count_launch();
  if (int rc = sb::check_launch("block_scan_kernel")) return rc;
  {
    sb::ProfScope prof("emit_kernel", st);
    emit_kernel<<<p.bd_blocks, BD_THREADS, 0, st>>>(codes, W, n, rin, perm_a, perm_b, ctrl, flags, counts, table_out,
                                                    reinterpret_cast<long long*>(row_code_out),
                                                    reinterpret_cast<long long*>(csr_off_out),
                                                    reinterpret_cast<long long*>(csr_rows_out));
    sb::count_launch();
    if (int rc = sb::check_launch("emit_kernel")) return rc;
  }
  {
    long long blocks = (n + 255) / 256;
    if (blocks > 4ll * sb::sm_count()) blocks = 4ll * sb::sm_count();
    max_count_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(csr_off_out),
                                                       reinterpret_cast<long long*>(stats_out));
    sb::count_launch();
    if (int rc = sb::check_launch("max_count_kernel")) return rc;
  }
  return SB_OK;
}

/* Sorted selection for LARGE candidate lists: same contract as sb_rerank_select_rows (cand_cnt / cand_idx
 * optional; cand_idx == NULL writes positions), but by one stable radix sort of (query, distance[, row])
 * keys over all M candidates -- O(M) instead of the per-query O(m^2) rank count. */
size_t sb_rerank_select_sorted_workspace_bytes(int64_t M) {
  if (!uc_supported(M, M, 4)) return 0;
  return al256((size_t)M * 16) + uc_plan(M, 4).total;
}

int sb_rerank_select_sorted(const double* dist, const int64_t* cand_off, const int64_t* cand_cnt, const int64_t* cand_idx,
                            int64_t M, int32_t Q, int32_t n, int32_t tie_by_row, int64_t* out_pos, double* out_dist,
                            void* workspace, size_t workspace_bytes, void* stream) {
  SB_REQUIRE(Q >= 1 && n >= 1 && M >= 1, "sb_rerank_select_sorted: bad sizes");
  SB_REQUIRE(dist && cand_off && out_pos && out_dist, "sb_rerank_select_sorted: NULL pointer");
  SB_REQUIRE(!tie_by_row || cand_idx, "sb_rerank_select_sorted: tie_by_row needs cand_idx");
  const size_t need = sb_rerank_select_sorted_workspace_bytes(M);
  if (need == 0) {
    sb::set_error("sb_rerank_select_sorted: M=%lld out of range", (long long)M);
    return SB_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < need) {
    sb::set_error("sb_rerank_select_sorted: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return SB_ERR_WORKSPACE;
  }
  SB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "sb_rerank_select_sorted: workspace must be 256-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint32_t* keys = static_cast<uint32_t*>(workspace);
  unsigned char* ws = static_cast<unsigned char*>(workspace) + al256((size_t)M * 16);
  const UcPlan p = uc_plan(M, 4);
  const long long* co = reinterpret_cast<const long long*>(cand_off);
  const long long* cc = reinterpret_cast<const long long*>(cand_cnt);
  const long long* ci = reinterpret_cast<const long long*>(cand_idx);
  {
    sb::ProfScope prof("select_keys_kernel", st);
    select_keys_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(dist, co, cc, ci, Q, M, tie_by_row ? 1 : 0, keys);
    sb::count_launch();
    if (int rc = sb::check_launch("select_keys_kernel")) return rc;
  }
  if (int rc = sort_rows_stage(keys, 4, nullptr, M, p, ws, st)) return rc;
  {
    sb::ProfScope prof("select_emit_kernel", st);
    const long long items = (long long)Q * n;
    select_emit_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(
        dist, co, cc, ci, Q, n, reinterpret_cast<uint32_t*>(ws + p.off_perm_a), reinterpret_cast<uint32_t*>(ws + p.off_perm_b),
        reinterpret_cast<SortCtrl*>(ws + p.off_ctrl), reinterpret_cast<long long*>(out_pos), out_dist);
    sb::count_launch();
    if (int rc = sb::check_launch("select_emit_kernel")) return rc;
  }
  return SB_OK;
}

/* Exhaustive Hamming top-k for k beyond the scan kernels' list size (the reference takes any n,
 * linear.py:232-240): every (query, row) key is materialised and sorted.  Q * U < 2^30 per call. */
size_t sb_hamming_topk_sorted_workspace_bytes(int64_t U, int32_t Q) {
  const int64_t M = U * (int64_t)Q;
  if (U < 1 || Q < 1 || !uc_supported(M, M, 4)) return 0;
  return al256((size_t)M * 16) + uc_plan(M, 4).total;
}

int sb_hamming_topk_sorted(const uint32_t* db, int64_t U, int32_t W, const uint32_t* q, int32_t Q, int32_t k, int64_t idx_base,
                           uint64_t* keys_out, void* workspace, size_t workspace_bytes, void* stream) {
  SB_REQUIRE(db && q && keys_out, "sb_hamming_topk_sorted: NULL pointer");
  SB_REQUIRE(k >= 1 && idx_base >= 0 && idx_base + U < (1ll << 40), "sb_hamming_topk_sorted: bad k / idx_base");
  SB_REQUIRE(W == 1 || W == 2 || W == 4 || W == 8 || W == 16 || W == 32, "sb_hamming_topk_sorted: W must be a power of two <= 32");
  const size_t need = sb_hamming_topk_sorted_workspace_bytes(U, Q);
  if (need == 0) {
    sb::set_error("sb_hamming_topk_sorted: Q * U = %lld out of range (< 2^30)", (long long)(U * (int64_t)Q));
    return SB_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < need) {
    sb::set_error("sb_hamming_topk_sorted: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return SB_ERR_WORKSPACE;
  }
  SB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "sb_hamming_topk_sorted: workspace must be 256-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t M = U * (int64_t)Q;
  uint32_t* keys = static_cast<uint32_t*>(workspace);
  unsigned char* ws = static_cast<unsigned char*>(workspace) + al256((size_t)M * 16);
  const UcPlan p = uc_plan(M, 4);
  {
    sb::ProfScope prof("ham_all_keys_kernel", st);
    const unsigned grid = (unsigned)((U + 255) / 256);
    switch (W) {
      case 1: ham_all_keys_kernel<1><<<grid, 256, 0, st>>>(db, U, q, Q, idx_base, keys); break;
      case 2: ham_all_keys_kernel<2><<<grid, 256, 0, st>>>(db, U, q, Q, idx_base, keys); break;
      case 4: ham_all_keys_kernel<4><<<grid, 256, 0, st>>>(db, U, q, Q, idx_base, keys); break;
      case 8: ham_all_keys_kernel<8><<<grid, 256, 0, st>>>(db, U, q, Q, idx_base, keys); break;
      case 16: ham_all_keys_kernel<16><<<grid, 256, 0, st>>>(db, U, q, Q, idx_base, keys); break;
      default: ham_all_keys_kernel<32><<<grid, 256, 0, st>>>(db, U, q, Q, idx_base, keys); break;
    }
    sb::count_launch();
    if (int rc = sb::check_launch("ham_all_keys_kernel")) return rc;
  }
  if (int rc = sort_rows_stage(keys, 4, nullptr, M, p, ws, st)) return rc;
  {
    sb::ProfScope prof("ham_all_emit_kernel", st);
    const long long items = (long long)Q * k;
    ham_all_emit_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(
        keys, U, Q, k, reinterpret_cast<uint32_t*>(ws + p.off_perm_a), reinterpret_cast<uint32_t*>(ws + p.off_perm_b),
        reinterpret_cast<SortCtrl*>(ws + p.off_ctrl), reinterpret_cast<unsigned long long*>(keys_out));
    sb::count_launch();
    if (int rc = sb::check_launch("ham_all_emit_kernel")) return rc;
  }
  return SB_OK;
}

}  // extern "C"
