// Thin inline-PTX wrappers shared by the tcgen05 kernels (itq_hash_tc.cu, flat_l2_tma.cu):
// mbarrier, TMA bulk / tensor copies, proxy fences, tcgen05.mma / commit / ld, TF32 rounding.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tcptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TC_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TC_WAIT_DONE;\n"
      "bra TC_WAIT_LOOP;\n"
      "TC_WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// streaming 128-bit load: X is read exactly once, keep it out of L1
__device__ __forceinline__ float4 ldg_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// Same rounding (nearest, ties away from zero) for FINITE inputs in two integer ops:
// cvt.rna.tf32.f32 compiles to add + Inf/NaN test + select + mask on sm_100a.
__device__ __forceinline__ uint32_t to_tf32_finite(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

// K-major, no-swizzle shared-memory matrix descriptor (tcgen05 "SmemDescriptor"):
// core matrix = 8 rows x 16 bytes stored as 128 contiguous bytes;
// LBO = byte distance between the two 16-byte K chunks of one MMA, SBO = byte
// distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (sm_100)
  return d;         // base_offset 0, layout_type 0 = SWIZZLE_NONE
}

// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32, issued by one thread.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// Same shape with .pack::16b: each register receives the low halves of TWO adjacent columns (column
// 2j in bits 0-15, 2j+1 in bits 16-31) -- 64 columns of 16-bit accumulators per call.
__device__ __forceinline__ void tmem_ld32_pack16_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// SWIZZLE_128B K-major descriptor: rows of 128 bytes, 8-row atoms of 1024 bytes (SBO), 16-byte
// chunks XOR-swizzled by (row % 8) -- what a TMA tensor copy with CU_TENSOR_MAP_SWIZZLE_128B
// writes.  The tile base must be 1024-byte aligned; K steps advance the start address.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;                      // LBO (unused for swizzled K-major layouts)
  d |= (uint64_t)(1024 >> 4) << 32;            // SBO: 8 rows x 128 bytes
  d |= 1ull << 46;                             // descriptor version (sm_100)
  d |= 2ull << 61;                             // layout_type = SWIZZLE_128B
  return d;
}

// 2-D tiled TMA load (tensor map in kernel parameter space), completes on an mbarrier.
__device__ __forceinline__ void tma_tensor_2d_g2s(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// ---- flat-L2 filter epilogue helpers (itq_hash_tc.cu EPI_L2, flat_l2_tma.cu) ----
// 16 accumulators -> 16-bit mask of the NON-NEGATIVE ones, one SHF per element: the sign bits are
// shifted into `neg` (element j lands on bit 15 - j), then inverted.
__device__ __forceinline__ unsigned nonneg_mask16(const uint32_t (&v)[16]) {
  unsigned neg = 0u;
#pragma unroll
  for (int j = 0; j < 16; ++j) neg = __funnelshift_l(v[j], neg, 1);
  return ~neg & 0xffffu;
}
// Cold path, deliberately out of line so that none of its address arithmetic is hoisted into the
// hot loop: append the pairs flagged in `mask` (bit 15 - j <-> column q0 + j) to their queries'
// candidate buffers.  key = (float bits of d2) << 32 | row, d2 = tq - 2 * accumulator.  The 16
// accumulators travel BY VALUE (registers): taking their address would pin the caller's arrays
// in local memory.
static __device__ __noinline__ void l2_append_survivors(
    unsigned mask, long long q0, unsigned row_id, const float* tq, unsigned long long* cand_buf, int* cand_cnt, int cap,
    uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3, uint32_t v4, uint32_t v5, uint32_t v6, uint32_t v7, uint32_t v8,
    uint32_t v9, uint32_t v10, uint32_t v11, uint32_t v12, uint32_t v13, uint32_t v14, uint32_t v15) {
  const uint32_t v[16] = {v0, v1, v2, v3, v4, v5, v6, v7, v8, v9, v10, v11, v12, v13, v14, v15};
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (mask & (0x8000u >> j)) {
      const long long qg = q0 + j;
      const float d2 = fmaxf(fmaf(-2.0f, __uint_as_float(v[j]), __ldcg(tq + qg)), 0.0f);
      const int slot = atomicAdd(cand_cnt + qg, 1);
      if (slot < cap) cand_buf[qg * cap + slot] = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)row_id;
    }
  }
}
#define SB_L2_APPEND(mask, v, q0, row_id, tq, buf, cnt, cap)                                                          \
  tcptx::l2_append_survivors(mask, q0, row_id, tq, buf, cnt, cap, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], \
                             v[9], v[10], v[11], v[12], v[13], v[14], v[15])

// Same cold path, but the survivors go to a per-warp shared-memory queue (drained after the
// accumulator buffer has been released); when the queue is full they are appended directly.
// Thresholds come from shared memory (`tq_local[col]`, this query block's 256 values): a dependent
// global load per survivor would cost more than everything else on this path.
static __device__ __noinline__ void l2_queue_survivors(
    unsigned mask, int q0, int col0, unsigned row_id, const float* tq_local, unsigned long long* q_key, int* q_q, int* q_cnt,
    int q_cap, unsigned long long* cand_buf, int* cand_cnt, int cap,
    uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3, uint32_t v4, uint32_t v5, uint32_t v6, uint32_t v7, uint32_t v8,
    uint32_t v9, uint32_t v10, uint32_t v11, uint32_t v12, uint32_t v13, uint32_t v14, uint32_t v15) {
  const uint32_t v[16] = {v0, v1, v2, v3, v4, v5, v6, v7, v8, v9, v10, v11, v12, v13, v14, v15};
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (mask & (0x8000u >> j)) {
      const int qg = q0 + col0 + j;
      const float d2 = fmaxf(fmaf(-2.0f, __uint_as_float(v[j]), tq_local[col0 + j]), 0.0f);
      const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)row_id;
      const int e = atomicAdd(q_cnt, 1);
      if (e < q_cap) {
        q_key[e] = key;
        q_q[e] = qg;
      } else {
        const int slot = atomicAdd(cand_cnt + qg, 1);
        if (slot < cap) cand_buf[(long long)qg * cap + slot] = key;
      }
    }
  }
}
#define SB_L2_QUEUE(mask, v, q0, col0, row_id, tql, qk, qq, qc, qcap, buf, cnt, cap)                                    \
  tcptx::l2_queue_survivors(mask, q0, col0, row_id, tql, qk, qq, qc, qcap, buf, cnt, cap, v[0], v[1], v[2], v[3], v[4],  \
                            v[5], v[6], v[7], v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15])

// Ask the TMA unit to pull a tile into L2 only (no shared-memory destination, no barrier).
__device__ __forceinline__ void tma_tensor_2d_prefetch_l2(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}

}  // namespace tcptx
