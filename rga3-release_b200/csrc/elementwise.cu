// Memory-bound row kernels: RMSNorm (fp32 residual stream -> bf16 GEMM operand) and
// the dtype cast in front of the patch-embed GEMM.  HBM-roofline kernels: one warp
// per row, 16-byte vector accesses, the row is held in registers between the two
// passes so x is read exactly once (4 B in, 2 B out per element).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "internal.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace b200 {
namespace {

// HF modeling_qwen2_5_vl.py:66-71: w * (x * rsqrt(mean(x^2) + eps)) with fp32 statistics.
// NV = float4 per lane (dim = 128 * NV), compile-time so all loads of a row are issued back to back;
// NV = 0 is the generic two-pass fallback.  Rows are distributed warp-by-warp over a grid sized to the
// machine (grid-stride), so the tail is one row, not one block.
template <int NV>
__global__ void __launch_bounds__(128)
rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int rows,
               int dim, float eps) {
  griddep_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const float4* wr = reinterpret_cast<const float4*>(w);
  const float inv_dim = 1.0f / static_cast<float>(dim);
  float4 g[NV > 0 ? NV : 1];
  if constexpr (NV > 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) g[i] = __ldg(wr + lane + i * 32);  // norm weights do not depend on the previous kernel
  }
  griddep_wait();
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * dim);
    uint2* orow = reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * dim);
    if constexpr (NV > 0) {
      float4 v[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = xr[lane + i * 32];
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rstd = rsqrtf(ss * inv_dim + eps);
#pragma unroll
      for (int i = 0; i < NV; ++i)
        orow[lane + i * 32] = make_uint2(pack_bf16x2(v[i].x * rstd * g[i].x, v[i].y * rstd * g[i].y),
                                         pack_bf16x2(v[i].z * rstd * g[i].z, v[i].w * rstd * g[i].w));
    } else {
      const int nv = dim >> 2;
      float ss = 0.f;
      for (int idx = lane; idx < nv; idx += 32) {
        const float4 t = xr[idx];
        ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rstd = rsqrtf(ss * inv_dim + eps);
      for (int idx = lane; idx < nv; idx += 32) {
        const float4 t = xr[idx];
        const float4 gg = __ldg(wr + idx);
        orow[idx] = make_uint2(pack_bf16x2(t.x * rstd * gg.x, t.y * rstd * gg.y), pack_bf16x2(t.z * rstd * gg.z, t.w * rstd * gg.w));
      }
    }
  }
}

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T>
__global__ void __launch_bounds__(256) cast_bf16_kernel(const T* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
  griddep_launch_dependents();
  griddep_wait();
  const int64_t i0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i0 + 8 <= n) {
    float f[8];
    if constexpr (sizeof(T) == 4) {
      const float4 a = *reinterpret_cast<const float4*>(in + i0), b = *reinterpret_cast<const float4*>(in + i0 + 4);
      f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
    } else {
      const uint4 raw = *reinterpret_cast<const uint4*>(in + i0);
      const T* h = reinterpret_cast<const T*>(&raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = to_f<T>(h[j]);
    }
    *reinterpret_cast<uint4*>(out + i0) =
        make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  } else {
    for (int64_t i = i0; i < n; ++i) out[i] = __float2bfloat16_rn(to_f<T>(in[i]));
  }
}

}  // namespace

int launch_rmsnorm(const float* x, const float* w, void* out_bf16, int rows, int dim, float eps, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (dim % 4) return fail(B200VIT_EINVAL, "rmsnorm: dim must be a multiple of 4");
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out_bf16)) & 15)
    return fail(B200VIT_EALIGN, "rmsnorm: pointers must be 16-byte aligned");
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    const char* e = getenv("B200VIT_RMS_BPS");
    blocks_per_sm = e ? atoi(e) : 8;
  }
  int grid = device_sm_count() * blocks_per_sm;  // 4 warps per block
  if (grid > (rows + 3) / 4) grid = (rows + 3) / 4;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (dim == 1280)
    B200_CUDA_OK(launch_kernel(rmsnorm_kernel<10>, dim3(grid), dim3(128), 0, stream, 1, x, w, o, rows, dim, eps));
  else
    B200_CUDA_OK(launch_kernel(rmsnorm_kernel<0>, dim3(grid), dim3(128), 0, stream, 1, x, w, o, rows, dim, eps));
  return 0;
}

int launch_cast_bf16(const void* in, int in_dtype, void* out, int64_t n, cudaStream_t stream) {
  if (n <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15)
    return fail(B200VIT_EALIGN, "cast: pointers must be 16-byte aligned");
  const int64_t threads = (n + 7) / 8;
  const int grid = static_cast<int>((threads + 255) / 256);
  if (in_dtype == 0)
    B200_CUDA_OK(launch_kernel(cast_bf16_kernel<float>, dim3(grid), dim3(256), 0, stream, 1, reinterpret_cast<const float*>(in),
                               reinterpret_cast<__nv_bfloat16*>(out), n));
  else if (in_dtype == 1)
    B200_CUDA_OK(launch_kernel(cast_bf16_kernel<__half>, dim3(grid), dim3(256), 0, stream, 1, reinterpret_cast<const __half*>(in),
                               reinterpret_cast<__nv_bfloat16*>(out), n));
  else if (in_dtype == 2)
    B200_CUDA_OK(cudaMemcpyAsync(out, in, n * 2, cudaMemcpyDeviceToDevice, stream));
  else
    return fail(B200VIT_EINVAL, "cast: unknown dtype");
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace b200

// ------------------------------------------------------------------ SM clock probe (measurement aid)
// One thread spins for `spin_ns` of wall time and reports (SM cycles, nanoseconds) over that interval: the SM clock
// the GPU actually runs at while other kernels execute, read without NVML (bench.py uses it for N > 1, where NVML
// reads inside an NCCL-coupled loop perturb the measurement).
namespace b200 {
namespace {
__global__ void clock_probe_kernel(unsigned long long* __restrict__ out, unsigned int spin_ns) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  const long long c0 = clock64();
  do {
    __nanosleep(200);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while (t1 - t0 < spin_ns);
  const long long c1 = clock64();
  out[0] = static_cast<unsigned long long>(c1 - c0);
  out[1] = t1 - t0;
}
}  // namespace
}  // namespace b200

extern "C" int b200vit_clock_probe(uint64_t* d_out2, uint32_t spin_ns, b200vit_stream stream) {
  if (d_out2 == nullptr || (reinterpret_cast<uintptr_t>(d_out2) & 7)) return b200::fail(B200VIT_EINVAL, "clock_probe: bad output pointer");
  if (spin_ns > 1000000u) spin_ns = 1000000u;
  b200::clock_probe_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned long long*>(d_out2), spin_ns);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

