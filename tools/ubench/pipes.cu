// Micro-benchmark: per-SM throughput of the instruction mix of the attention softmax (MUFU.EX2, F2FP, FFMA/FADD)
// and of a polynomial exp2 on the FMA pipe.  nvcc -arch=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
// 2^x for x <= 0 on the FMA pipe: split into integer and fraction, degree-3 minimax on [0,1), exponent add
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  float fl = floorf(x);   // FRND (may be a slower pipe) -- alternative below
  float f = x - fl;
  float p = fmaf(f, 0.0555054f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (static_cast<int>(fl) << 23));
}
// magic-number variant: no FRND/F2I
__device__ __forceinline__ float ex2_poly2(float x) {
  x = fmaxf(x, -126.f);
  const float magic = 12582912.f;  // 1.5 * 2^23
  float r = x + magic;              // round to nearest integer in the low mantissa bits
  float fl = r - magic;
  float f = x - fl;                 // in [-0.5, 0.5]
  float p = fmaf(f, 0.0555054f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float seed) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed * (threadIdx.x + i);
  float acc = 0.f;
  uint32_t acc_u = 0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      if (MODE == 0) { v[i] = ex2(v[i]); v[i + 1] = ex2(v[i + 1]); }
      if (MODE == 1) { acc_u ^= pack(v[i], v[i + 1]); v[i] += 1.f; }
      if (MODE == 2) {  // softmax pass-2 mix: ffma, ex2, fadd, half a pack
        float p0 = ex2(fmaf(v[i], seed, -1.f)), p1 = ex2(fmaf(v[i + 1], seed, -1.f));
        acc += p0; acc += p1; acc_u ^= pack(p0, p1); v[i] = p0; v[i + 1] = p1;
      }
      if (MODE == 3) { v[i] = ex2_poly(v[i] - 3.f); v[i + 1] = ex2_poly(v[i + 1] - 3.f); }
      if (MODE == 4) { v[i] = ex2_poly2(v[i] - 3.f); v[i + 1] = ex2_poly2(v[i + 1] - 3.f); }
      if (MODE == 5) {  // mix with every 4th exponential on the FMA pipe
        float a0 = fmaf(v[i], seed, -1.f), a1 = fmaf(v[i + 1], seed, -1.f);
        float p0 = (i == 0) ? ex2_poly2(a0) : ex2(a0), p1 = (i == 0) ? ex2_poly2(a1) : ex2(a1);
        acc += p0; acc += p1; acc_u ^= pack(p0, p1); v[i] = p0; v[i + 1] = p1;
      }
      if (MODE == 6) { v[i] = fmaf(v[i], seed, 1.f); v[i + 1] = fmaf(v[i + 1], seed, 1.f); }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = acc + __uint_as_float(acc_u);
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<MODE><<<148, threads>>>(out, cyc, iters, 0.001f);
  k<MODE><<<148, threads>>>(out, cyc, iters, 0.001f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = h[0];
  printf("%-28s threads %4d: %.2f elements/clk/SM  (%.2f clk per warp-element)\n", name, threads, double(iters) * 8 * threads / c,
         c / (double(iters) * 8 * threads / 32));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int th : {128, 256, 512}) {
    run<0>("ex2", th);
    run<1>("cvt.bf16x2 (pairs)", th);
    run<2>("softmax mix", th);
    run<3>("poly exp2 (floor)", th);
    run<4>("poly exp2 (magic)", th);
    run<5>("mix, 1/4 poly", th);
    run<6>("ffma", th);
  }
  // accuracy of the polynomial
  return 0;
}
