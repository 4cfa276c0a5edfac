// Memory-bound row kernels: RMSNorm (fp32 residual stream -> bf16 GEMM operand) and
// the dtype cast in front of the patch-embed GEMM.  HBM-roofline kernels: one warp
// per row, 16-byte vector accesses, the row is held in registers between the two
// passes so x is read exactly once (4 B in, 2 B out per element).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "internal.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace b200 {
namespace {

constexpr int RMS_MAX_V4 = 16;  // float4 per lane held in registers: rows up to 2048 floats

// HF modeling_qwen2_5_vl.py:66-71: w * (x * rsqrt(mean(x^2) + eps)) with fp32 statistics.
__global__ void __launch_bounds__(256)
rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int rows,
               int dim, float eps) {
  griddep_launch_dependents();
  griddep_wait();
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp_global >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(warp_global) * dim);
  const float4* wr = reinterpret_cast<const float4*>(w);
  const int nv = dim >> 2;
  float4 v[RMS_MAX_V4];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < RMS_MAX_V4; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      v[i] = xr[idx];
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
  for (int idx = lane + RMS_MAX_V4 * 32; idx < nv; idx += 32) {  // rows longer than the register cache
    float4 t = xr[idx];
    ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / static_cast<float>(dim) + eps);
  uint2* orow = reinterpret_cast<uint2*>(out + static_cast<size_t>(warp_global) * dim);
#pragma unroll
  for (int i = 0; i < RMS_MAX_V4; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      const float4 g = __ldg(wr + idx);
      orow[idx] = make_uint2(pack_bf16x2(v[i].x * rstd * g.x, v[i].y * rstd * g.y),
                             pack_bf16x2(v[i].z * rstd * g.z, v[i].w * rstd * g.w));
    }
  }
  for (int idx = lane + RMS_MAX_V4 * 32; idx < nv; idx += 32) {
    const float4 t = xr[idx];
    const float4 g = __ldg(wr + idx);
    orow[idx] = make_uint2(pack_bf16x2(t.x * rstd * g.x, t.y * rstd * g.y), pack_bf16x2(t.z * rstd * g.z, t.w * rstd * g.w));
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
  const int warps_per_block = 8;
  const int grid = (rows + warps_per_block - 1) / warps_per_block;
  B200_CUDA_OK(launch_kernel(rmsnorm_kernel, dim3(grid), dim3(warps_per_block * 32), 0, stream, 1, x, w,
                             reinterpret_cast<__nv_bfloat16*>(out_bf16), rows, dim, eps));
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
