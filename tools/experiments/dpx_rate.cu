// Issue rate of the instructions of the FP4 scan epilogue on one SM sub-partition: VIMNMX3.U16x2, LOP3, SHF, FADD.RZ
// and the epilogue's own mix.  One CTA, `warps` warps; every thread runs 8 independent dependency chains.
// Measured on a B200: VIMNMX3.U16x2 0.50 / clk / SMSP (= SHF), LOP3 + FADD.RZ 0.93, the epilogue's mix 1.08
// (8 instructions per 2 accumulators -> 3.7 clk per accumulator and sub-partition).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o dpx_rate dpx_rate.cu ; run: ./dpx_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void rate(uint32_t* out, int iters, long long* cycles, uint32_t seed) {
  uint32_t a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = seed * (threadIdx.x + j + 1); b[j] = seed ^ (j * 0x9e3779b9u + threadIdx.x); }
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (OP == 0) a[j] = __vimax3_u16x2(a[j], b[j], b[(j + 1) & 7]);
      if (OP == 1) a[j] = (a[j] & b[j]) ^ b[(j + 1) & 7];                                   // LOP3
      if (OP == 2) a[j] = __funnelshift_l(a[j], b[j], 9);                                    // SHF
      if (OP == 3) a[j] = __float_as_uint(__fadd_rz(__uint_as_float(a[j] & 0x3fffffffu), 25165824.0f));   // LOP3 + FADD.RZ
      if (OP == 4) {                                                                         // the epilogue's mix for 2 values
        const uint32_t v0 = __float_as_uint(__fadd_rz(__uint_as_float(b[j]), 25165824.0f));
        const uint32_t v1 = __float_as_uint(__fadd_rz(__uint_as_float(b[(j + 1) & 7]), 25165824.0f));
        a[j] = __vimin3_u16x2(a[j], v0 << 1, v1 << 1);
        b[j] = __vimax3_u16x2(b[j] | 1u, v0 << 9, v1 << 9);
      }
      if (OP == 5) a[j] = max(a[j], b[j]);                                                   // IMNMX
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s ^= a[j] ^ b[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
  for (int warps : {4, 16, 32}) {
    const int iters = 20000;
    rate<OP><<<1, warps * 32>>>(out, iters, cyc, 12345u);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double inst = (double)iters * 8 * per_iter * warps / 4.0;      // warp instructions per sub-partition
    printf("%-28s warps=%2d  %.3f warp-instr/clk/SMSP (%.2f clk per instr)\n", name, warps, inst / c, c / inst);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  // (plain IMNMX / LOP3 chains are folded by the compiler -- idempotent / algebraically reducible -- and report
  // impossible rates; SHF is the reference point for "one ALU-pipe instruction")
  run<0>("VIMNMX3.U16x2", 1);
  run<2>("SHF", 1);
  run<3>("LOP3+FADD.RZ", 2);
  run<4>("epilogue mix (8 per 2 values)", 8);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
