// Throughput of legacy warp-level mma.sync.m16n8k16 (bf16) on one SM of a B200: cycles per instruction per warp with
// W resident warps, 8 independent accumulator chains each.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 hmma.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, long long* cyc, int iters) {
  float d[8][4] = {};
  unsigned a0 = threadIdx.x, a1 = 2, a2 = 3, a3 = 4, b0 = 5, b1 = 6;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < 8; ++j) s += d[j][0] + d[j][1] + d[j][2] + d[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const int iters = 2000;
  for (int warps : {1, 2, 4, 8, 12, 16}) {
    k<<<1, 32 * warps>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    k<<<1, 32 * warps>>>(out, cyc, iters);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per = double(c) / (iters * 8);
    printf("warps %2d: %.1f cycles per HMMA per warp -> %.0f dense bf16 FMA/clk/SM\n", warps, per, warps * 4096.0 / per);
  }
  return 0;
}
