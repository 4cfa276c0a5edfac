// b200vit_pack_weights: HF state_dict tensors -> the packed device layout the kernels read.
//
// What the host module used to do in torch (and what a non-Python host would otherwise have to
// re-implement): bf16 casts, the RMSNorm gamma fold (HF modeling_qwen2_5_vl.py:57-71 -- RMSNorm(x) W^T ==
// rstd(x) * (x (W diag(gamma))^T), so norm1.weight scales the columns of attn.qkv.weight and norm2.weight those of
// mlp.gate_proj / up_proj), the gate/up row interleave the SwiGLU epilogue expects, and the I -> Ipad zero padding
// (3420 * 2 B rows are not 16-byte aligned, TMA-illegal).  One pass over the weights, HBM-bound, runs once.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstring>

#include "internal.h"

namespace b200 {
namespace {

size_t up256(size_t v) { return (v + 255) / 256 * 256; }

template <typename T>
__device__ __forceinline__ float as_f32(T v);
template <>
__device__ __forceinline__ float as_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float as_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float as_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// dst[(r * row_mul + row_off) * dst_cols + c] = bf16(src[r, c] * gamma[c]); dst is zero-filled beforehand
template <typename T>
__global__ void __launch_bounds__(256)
pack_matrix_kernel(const T* __restrict__ src, int64_t rows, int64_t cols, const T* __restrict__ gamma,
                   __nv_bfloat16* __restrict__ dst, int64_t dst_cols, int row_mul, int row_off) {
  const int64_t n = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    float v = as_f32<T>(src[i]);
    if (gamma != nullptr) v *= as_f32<T>(gamma[c]);  // fp32 product, one rounding to bf16
    dst[(r * row_mul + row_off) * dst_cols + c] = __float2bfloat16_rn(v);
  }
}

// qkv projection: HF rows are type-major [q heads | k heads | v heads]; the kernels want them head-major
// [(head, {q, k, v}, head_dim)] so that one 240-column GEMM tile holds Q_h | K_h | V_h of one head (the fused
// window-attention epilogue needs all three in one accumulator).  `cols` = 0 packs the bias vector the same way.
template <typename T>
__global__ void __launch_bounds__(256)
pack_qkv_kernel(const T* __restrict__ src, int64_t d_model, int64_t cols, int head_dim, const T* __restrict__ gamma,
                __nv_bfloat16* __restrict__ dst_w, float* __restrict__ dst_b) {
  const int64_t width = cols > 0 ? cols : 1;
  const int64_t n = 3 * d_model * width;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / width, c = i - r * width;          // destination row: (h * 3 + t) * head_dim + d
    const int64_t h = r / (3 * head_dim), t = (r / head_dim) % 3, d = r % head_dim;
    const int64_t sr = t * d_model + h * head_dim + d;
    if (cols > 0) {
      float v = as_f32<T>(src[sr * cols + c]);
      if (gamma != nullptr) v *= as_f32<T>(gamma[c]);
      dst_w[i] = __float2bfloat16_rn(v);
    } else {
      dst_b[r] = as_f32<T>(src[sr]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
pack_vector_kernel(const T* __restrict__ src, int64_t n, float* __restrict__ dst, int mul, int off) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i * mul + off] = as_f32<T>(src[i]);
}

int elem_bytes(int dtype) { return dtype == 0 ? 4 : 2; }

// Device view of a tensor that may live on the host: host tensors are staged into a scratch buffer.
struct Stager {
  cudaStream_t stream;
  int dtype;
  void* scratch[2] = {nullptr, nullptr};
  size_t cap[2] = {0, 0};
  bool dirty[2] = {false, false};
  ~Stager() {
    cudaFree(scratch[0]);
    cudaFree(scratch[1]);
  }
  int view(const void* p, size_t elems, int slot, const void** out) {
    cudaPointerAttributes at;
    const cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) cudaGetLastError();
    if (e == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)) {
      *out = p;
      return 0;
    }
    const size_t bytes = elems * elem_bytes(dtype);
    if (dirty[slot]) {  // the kernel that read the previous contents must have finished
      B200_CUDA_OK(cudaStreamSynchronize(stream));
      dirty[0] = dirty[1] = false;
    }
    if (cap[slot] < bytes) {
      cudaFree(scratch[slot]);
      scratch[slot] = nullptr;
      B200_CUDA_OK(cudaMalloc(&scratch[slot], bytes));
      cap[slot] = bytes;
    }
    B200_CUDA_OK(cudaMemcpy(scratch[slot], p, bytes, cudaMemcpyHostToDevice));
    dirty[slot] = true;
    *out = scratch[slot];
    return 0;
  }
};

struct Packer {
  Stager st;
  int matrix(const void* src, int64_t rows, int64_t cols, const void* gamma, void* dst, int64_t dst_cols, int row_mul, int row_off) {
    if (!src) return fail(B200VIT_EINVAL, "pack_weights: null tensor");
    const void *s = nullptr, *g = nullptr;
    int rc;
    if ((rc = st.view(src, rows * cols, 0, &s))) return rc;
    if (gamma && (rc = st.view(gamma, cols, 1, &g))) return rc;
    const int64_t n = rows * cols;
    const int grid = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 16));
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
    if (st.dtype == 0)
      pack_matrix_kernel<float><<<grid, 256, 0, st.stream>>>(reinterpret_cast<const float*>(s), rows, cols, reinterpret_cast<const float*>(g), d, dst_cols, row_mul, row_off);
    else if (st.dtype == 1)
      pack_matrix_kernel<__half><<<grid, 256, 0, st.stream>>>(reinterpret_cast<const __half*>(s), rows, cols, reinterpret_cast<const __half*>(g), d, dst_cols, row_mul, row_off);
    else
      pack_matrix_kernel<__nv_bfloat16><<<grid, 256, 0, st.stream>>>(reinterpret_cast<const __nv_bfloat16*>(s), rows, cols, reinterpret_cast<const __nv_bfloat16*>(g), d, dst_cols, row_mul, row_off);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
  }
  // head-interleaved qkv weight ([3D, D] with gamma folded) or, with cols = 0, bias ([3D])
  int qkv(const void* src, int64_t d_model, int64_t cols, int head_dim, const void* gamma, void* dst) {
    if (!src) return fail(B200VIT_EINVAL, "pack_weights: null tensor");
    const void *s = nullptr, *g = nullptr;
    int rc;
    if ((rc = st.view(src, 3 * d_model * (cols > 0 ? cols : 1), 0, &s))) return rc;
    if (gamma && (rc = st.view(gamma, cols, 1, &g))) return rc;
    const int64_t n = 3 * d_model * (cols > 0 ? cols : 1);
    const int grid = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 16));
    __nv_bfloat16* dw = reinterpret_cast<__nv_bfloat16*>(dst);
    float* db = reinterpret_cast<float*>(dst);
    if (st.dtype == 0)
      pack_qkv_kernel<float><<<grid, 256, 0, st.stream>>>(reinterpret_cast<const float*>(s), d_model, cols, head_dim, reinterpret_cast<const float*>(g), dw, db);
    else if (st.dtype == 1)
      pack_qkv_kernel<__half><<<grid, 256, 0, st.stream>>>(reinterpret_cast<const __half*>(s), d_model, cols, head_dim, reinterpret_cast<const __half*>(g), dw, db);
    else
      pack_qkv_kernel<__nv_bfloat16><<<grid, 256, 0, st.stream>>>(reinterpret_cast<const __nv_bfloat16*>(s), d_model, cols, head_dim, reinterpret_cast<const __nv_bfloat16*>(g), dw, db);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
  }
  int vector(const void* src, int64_t n, void* dst, int mul, int off) {
    if (!src) return fail(B200VIT_EINVAL, "pack_weights: null tensor");
    const void* s = nullptr;
    int rc;
    if ((rc = st.view(src, n, 0, &s))) return rc;
    const int grid = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 16));
    float* d = reinterpret_cast<float*>(dst);
    if (st.dtype == 0) pack_vector_kernel<float><<<grid, 256, 0, st.stream>>>(reinterpret_cast<const float*>(s), n, d, mul, off);
    else if (st.dtype == 1) pack_vector_kernel<__half><<<grid, 256, 0, st.stream>>>(reinterpret_cast<const __half*>(s), n, d, mul, off);
    else pack_vector_kernel<__nv_bfloat16><<<grid, 256, 0, st.stream>>>(reinterpret_cast<const __nv_bfloat16*>(s), n, d, mul, off);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
  }
};

struct Layout {
  size_t d, i, ipad, o, kpe, u;
  explicit Layout(const b200vit_cfg& c)
      : d(c.hidden), i(c.intermediate), ipad((c.intermediate + 127) / 128 * 128), o(c.out_hidden),
        kpe(static_cast<size_t>(c.in_channels) * c.temporal_patch * c.patch * c.patch), u(static_cast<size_t>(c.merge) * c.merge) {}
  size_t layer_bytes() const {
    return up256(3 * d * d * 2) + up256(3 * d * 4) + up256(d * d * 2) + up256(d * 4) + up256(2 * ipad * d * 2) +
           up256(2 * ipad * 4) + up256(d * ipad * 2) + up256(d * 4);
  }
  size_t total(int depth) const {
    return up256(d * kpe * 2) + depth * layer_bytes() + up256(d * 4) + up256(u * d * u * d * 2) + up256(u * d * 4) +
           up256(o * u * d * 2) + up256(o * 4);
  }
};

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" size_t b200vit_packed_weights_bytes(const b200vit_cfg* cfg) {
  if (!cfg || cfg->depth <= 0) return 0;
  return Layout(*cfg).total(cfg->depth);
}

extern "C" int b200vit_pack_weights(const b200vit_cfg* cfg, const b200vit_raw_weights* raw, void* d_packed, size_t packed_bytes,
                                    b200vit_weights* out, b200vit_layer_weights* out_layers, b200vit_stream stream_) {
  if (!cfg || !raw || !d_packed || !out || !out_layers || !raw->layers) return fail(B200VIT_EINVAL, "pack_weights: null argument");
  if (raw->dtype < 0 || raw->dtype > 2) return fail(B200VIT_EINVAL, "pack_weights: dtype must be 0 (fp32), 1 (fp16) or 2 (bf16)");
  if (cfg->depth <= 0 || cfg->hidden <= 0 || cfg->intermediate <= 0) return fail(B200VIT_EINVAL, "pack_weights: bad config");
  const Layout L(*cfg);
  if (packed_bytes < L.total(cfg->depth)) return fail(B200VIT_ENOMEM, "pack_weights: buffer smaller than b200vit_packed_weights_bytes()");
  if (reinterpret_cast<uintptr_t>(d_packed) & 255) return fail(B200VIT_EALIGN, "pack_weights: buffer must be 256-byte aligned");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_CUDA_OK(cudaMemsetAsync(d_packed, 0, L.total(cfg->depth), stream));  // padding rows / columns stay zero
  Packer pk;
  pk.st.stream = stream;
  pk.st.dtype = raw->dtype;
  uint8_t* cur = reinterpret_cast<uint8_t*>(d_packed);
  auto take = [&](size_t bytes) {
    uint8_t* p = cur;
    cur += up256(bytes);
    return p;
  };
  const int64_t d = L.d, i = L.i, ipad = L.ipad, o = L.o, ud = L.u * L.d;
  if (cfg->heads <= 0 || cfg->hidden % cfg->heads) return fail(B200VIT_EINVAL, "pack_weights: bad hidden/heads");
  const int head_dim = cfg->hidden / cfg->heads;
  int rc;
  void* p;
  p = take(L.d * L.kpe * 2);
  if ((rc = pk.matrix(raw->patch_w, d, L.kpe, nullptr, p, L.kpe, 1, 0))) return rc;
  out->patch_w = p;
  for (int l = 0; l < cfg->depth; ++l) {
    const b200vit_raw_layer& r = raw->layers[l];
    b200vit_layer_weights& w = out_layers[l];
    p = take(3 * L.d * L.d * 2);
    if ((rc = pk.qkv(r.qkv_w, d, d, head_dim, r.norm1_w, p))) return rc;
    w.qkv_w = p;
    p = take(3 * L.d * 4);
    if ((rc = pk.qkv(r.qkv_b, d, 0, head_dim, nullptr, p))) return rc;
    w.qkv_b = reinterpret_cast<const float*>(p);
    p = take(L.d * L.d * 2);
    if ((rc = pk.matrix(r.proj_w, d, d, nullptr, p, d, 1, 0))) return rc;
    w.proj_w = p;
    p = take(L.d * 4);
    if ((rc = pk.vector(r.proj_b, d, p, 1, 0))) return rc;
    w.proj_b = reinterpret_cast<const float*>(p);
    p = take(2 * L.ipad * L.d * 2);
    if ((rc = pk.matrix(r.gate_w, i, d, r.norm2_w, p, d, 2, 0))) return rc;
    if ((rc = pk.matrix(r.up_w, i, d, r.norm2_w, p, d, 2, 1))) return rc;
    w.gateup_w = p;
    p = take(2 * L.ipad * 4);
    if ((rc = pk.vector(r.gate_b, i, p, 2, 0))) return rc;
    if ((rc = pk.vector(r.up_b, i, p, 2, 1))) return rc;
    w.gateup_b = reinterpret_cast<const float*>(p);
    p = take(L.d * L.ipad * 2);
    if ((rc = pk.matrix(r.down_w, d, i, nullptr, p, ipad, 1, 0))) return rc;
    w.down_w = p;
    p = take(L.d * 4);
    if ((rc = pk.vector(r.down_b, d, p, 1, 0))) return rc;
    w.down_b = reinterpret_cast<const float*>(p);
  }
  p = take(L.d * 4);
  if ((rc = pk.vector(raw->merger_ln_w, d, p, 1, 0))) return rc;
  out->merger_ln_w = reinterpret_cast<const float*>(p);
  p = take(ud * ud * 2);
  if ((rc = pk.matrix(raw->merger_fc1_w, ud, ud, nullptr, p, ud, 1, 0))) return rc;
  out->merger_fc1_w = p;
  p = take(ud * 4);
  if ((rc = pk.vector(raw->merger_fc1_b, ud, p, 1, 0))) return rc;
  out->merger_fc1_b = reinterpret_cast<const float*>(p);
  p = take(o * ud * 2);
  if ((rc = pk.matrix(raw->merger_fc2_w, o, ud, nullptr, p, ud, 1, 0))) return rc;
  out->merger_fc2_w = p;
  p = take(o * 4);
  if ((rc = pk.vector(raw->merger_fc2_b, o, p, 1, 0))) return rc;
  out->merger_fc2_b = reinterpret_cast<const float*>(p);
  out->layers = out_layers;
  out->ipad = static_cast<int32_t>(ipad);
  if (pk.st.dirty[0] || pk.st.dirty[1]) B200_CUDA_OK(cudaStreamSynchronize(stream));  // staged copies are freed on return
  return 0;
}
