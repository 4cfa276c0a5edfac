// Varlen (cu_seqlens-driven) non-causal attention for head_dim 80, bf16 in / bf16 out,
// fp32 softmax statistics and accumulation.  One kernel serves both the 64-patch
// window layers and the per-temporal-slice "full" layers of the Qwen2.5-VL tower
// (HF modeling_qwen2_5_vl.py:244-283, softmax in fp32 :199): a CTA takes <= 64 query
// rows of one segment and streams that segment's K/V in 64-row blocks (flash style,
// online softmax).  Q and K arrive with RoPE already applied by the QKV GEMM epilogue.
//
// v1 math path: mma.sync m16n8k16 (ldmatrix-fed).  Attention is 2.3 % of the tower's
// FLOPs; the tcgen05 version is the follow-up.
#include <cuda_bf16.h>

#include "internal.h"
#include "ptx.cuh"

namespace b200 {
namespace {

constexpr int HD = 80;    // head dim
constexpr int QB = 64;    // query rows per CTA
constexpr int KVB = 64;   // kv rows per block
constexpr int LDS = 88;   // padded smem row (elements): 176 B rows -> conflict-free ldmatrix
constexpr int TILE_ELEMS = 64 * LDS;
constexpr int SMEM_BYTES = 5 * TILE_ELEMS * 2;  // Q + 2x(K,V)

__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, int ld, int rows_valid) {
  // 64 rows x 10 chunks of 16 B; rows >= rows_valid are zero-filled
  for (int i = threadIdx.x; i < 64 * 10; i += 128) {
    const int r = i / 10, c = i % 10;
    const bool ok = r < rows_valid;
    const __nv_bfloat16* src = ok ? g + static_cast<size_t>(r) * ld + c * 8 : g;
    cp_async16(s + r * LDS + c * 8, src, ok ? 16u : 0u);
  }
}

__global__ void __launch_bounds__(128)
attn_varlen_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                   const AttnWork* __restrict__ work, int heads, float scale_log2) {
  griddep_launch_dependents();
  griddep_wait();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sK = sQ + TILE_ELEMS;       // [2][64][LDS]
  __nv_bfloat16* sV = sK + 2 * TILE_ELEMS;   // [2][64][LDS]

  const AttnWork w = work[blockIdx.x];
  const int head = blockIdx.y;
  const int D = heads * HD;
  const int ld = 3 * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;

  const __nv_bfloat16* qg = qkv + static_cast<size_t>(w.q_start) * ld + head * HD;
  const __nv_bfloat16* kg = qkv + static_cast<size_t>(w.kv_start) * ld + D + head * HD;
  const __nv_bfloat16* vg = kg + D;
  const int nblk = (w.kv_len + KVB - 1) / KVB;

  load_tile(sQ, qg, ld, w.q_len);
  load_tile(sK, kg, ld, min(KVB, w.kv_len));
  load_tile(sV, vg, ld, min(KVB, w.kv_len));
  cp_async_commit();

  uint32_t qa[5][4];
  float o[10][4];
#pragma unroll
  for (int i = 0; i < 10; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const bool active = warp * 16 < w.q_len;

  for (int j = 0; j < nblk; ++j) {
    const int buf = j & 1;
    if (j + 1 < nblk) {
      const int rows = min(KVB, w.kv_len - (j + 1) * KVB);
      load_tile(sK + (buf ^ 1) * TILE_ELEMS, kg + static_cast<size_t>(j + 1) * KVB * ld, ld, rows);
      load_tile(sV + (buf ^ 1) * TILE_ELEMS, vg + static_cast<size_t>(j + 1) * KVB * ld, ld, rows);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (active) {
      if (j == 0) {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) {
          const int mi = lane >> 3;
          const int r = warp * 16 + (mi & 1) * 8 + (lane & 7);
          const int c = ks * 16 + (mi >> 1) * 8;
          ldmatrix_x4(smem_u32(sQ + r * LDS + c), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
        }
      }
      const __nv_bfloat16* k_s = sK + buf * TILE_ELEMS;
      const __nv_bfloat16* v_s = sV + buf * TILE_ELEMS;
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      // S = Q K^T
#pragma unroll
      for (int np = 0; np < 4; ++np) {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) {
          const int mi = lane >> 3;
          const int r = (2 * np + (mi >> 1)) * 8 + (lane & 7);
          const int c = ks * 16 + (mi & 1) * 8;
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(smem_u32(k_s + r * LDS + c), b0, b1, b2, b3);
          mma_bf16_16816(s[2 * np], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b0, b1);
          mma_bf16_16816(s[2 * np + 1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b2, b3);
        }
      }
      // mask the ragged tail of the segment
      const int kv_left = w.kv_len - j * KVB;
      if (kv_left < KVB) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int c = nt * 8 + tq * 2;
          if (c >= kv_left) s[nt][0] = s[nt][2] = -INFINITY;
          if (c + 1 >= kv_left) s[nt][1] = s[nt][3] = -INFINITY;
        }
      }
      // online softmax (base-2, scale folded in)
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0 * scale_log2), mn1 = fmaxf(m1, mx1 * scale_log2);
      const float a0 = exp2f(m0 - mn0), a1 = exp2f(m1 - mn1);
      m0 = mn0;
      m1 = mn1;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = exp2f(s[nt][0] * scale_log2 - mn0);
        s[nt][1] = exp2f(s[nt][1] * scale_log2 - mn0);
        s[nt][2] = exp2f(s[nt][2] * scale_log2 - mn1);
        s[nt][3] = exp2f(s[nt][3] * scale_log2 - mn1);
        ps0 += s[nt][0] + s[nt][1];
        ps1 += s[nt][2] + s[nt][3];
      }
      l0 = l0 * a0 + ps0;
      l1 = l1 * a1 + ps1;
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        o[i][0] *= a0, o[i][1] *= a0, o[i][2] *= a1, o[i][3] *= a1;
      }
      // O += P V
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t pa0 = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        const uint32_t pa1 = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        const uint32_t pa2 = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        const uint32_t pa3 = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dp = 0; dp < 5; ++dp) {
          const int mi = lane >> 3;
          const int r = kk * 16 + (mi & 1) * 8 + (lane & 7);
          const int c = (2 * dp + (mi >> 1)) * 8;
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4_trans(smem_u32(v_s + r * LDS + c), b0, b1, b2, b3);
          mma_bf16_16816(o[2 * dp], pa0, pa1, pa2, pa3, b0, b1);
          mma_bf16_16816(o[2 * dp + 1], pa0, pa1, pa2, pa3, b2, b3);
        }
      }
    }
    __syncthreads();  // everyone is done with buf before it is refilled
  }

  if (active) {
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    __nv_bfloat16* og = out + static_cast<size_t>(w.q_start) * D + head * HD + tq * 2;
#pragma unroll
    for (int nt = 0; nt < 10; ++nt) {
      if (r0 < w.q_len)
        *reinterpret_cast<uint32_t*>(og + static_cast<size_t>(r0) * D + nt * 8) = pack_bf16x2(o[nt][0] * i0, o[nt][1] * i0);
      if (r1 < w.q_len)
        *reinterpret_cast<uint32_t*>(og + static_cast<size_t>(r1) * D + nt * 8) = pack_bf16x2(o[nt][2] * i1, o[nt][3] * i1);
    }
  }
}

}  // namespace

int launch_attention(const void* qkv, void* out, const AttnWork* d_work, int n_work, int heads, cudaStream_t stream) {
  if (n_work <= 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(attn_varlen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  dim3 grid(n_work, heads);
  attn_varlen_kernel<<<grid, 128, SMEM_BYTES, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                         reinterpret_cast<__nv_bfloat16*>(out), d_work, heads, scale_log2);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace b200
