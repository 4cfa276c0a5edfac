// tcgen05 / TMEM / TMA bf16 GEMM with fused epilogues for the vision tower.
//
//   out = epilogue( A[M,K] * B[N,K]^T ),  A, B bf16 K-major, fp32 accumulation in TMEM.
//
// Replaces the cuBLAS calls behind nn.Linear / Conv3d in the HF tower
// (modeling_qwen2_5_vl.py: patch_embed :106-114, qkv :214/:231, proj :215/:286,
// gate/up/down :82-88, merger :133-146) together with the elementwise work that
// follows each of them (2-D RoPE :149-167, residual adds :313/:320, SiLU-gate :88,
// exact GELU :139, window reorder :478-481, un-reorder :512-513).
//
// Structure: persistent CTAs (one per SM), warp-specialised:
//   warp 0      TMA producer   (A and B tiles, 128-B swizzle, STAGES-deep mbarrier ring)
//   warp 1      MMA issuer     (one thread, tcgen05.mma 128 x BN x 16, two TMEM accumulators)
//   warp 2      TMEM allocator
//   warps 4..   epilogue       (EG groups of 4 warps; tcgen05.ld -> registers -> global)
// so the epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda_bf16.h>

#include "internal.h"
#include "ptx.cuh"

namespace b200 {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle atom row

struct GemmParams {
  void* out;
  const float* bias;
  const int32_t* row_map;
  const float* cos;
  const float* sin;
  int m, n, k, ldo, rope_cols;
};

template <int BN>
struct TileCfg {
  static constexpr int STAGES = (BN > 128) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ACC_STRIDE = (BN <= 128) ? 128 : 256;  // TMEM columns between the two accumulators
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + slack for 1024-B alignment
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------ epilogues
// Each epilogue thread owns one accumulator row (TMEM lane) and CW consecutive columns.
template <int EPI, int CW>
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, int row, int col0, const GemmParams& p) {
  const bool row_ok = row < p.m;
  if constexpr (EPI == B200VIT_EPI_QKV_ROPE) {
    static_assert(CW == 80, "QKV epilogue handles one 80-wide head per warp group");
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + col0;
    if (col0 < p.rope_cols) {
      const float* cs = p.cos + static_cast<size_t>(row) * 40;
      const float* sn = p.sin + static_cast<size_t>(row) * 40;
#pragma unroll
      for (int d0 = 0; d0 < 40; d0 += 8) {
        uint32_t lo[8], hi[8];
        tmem_ld8(taddr + d0, lo);
        tmem_ld8(taddr + 40 + d0, hi);
        tmem_ld_wait();
        if (row_ok) {
          float c[8], s[8], bl[8], bh[8];
          *reinterpret_cast<float4*>(&c[0]) = ldg4(cs + d0);
          *reinterpret_cast<float4*>(&c[4]) = ldg4(cs + d0 + 4);
          *reinterpret_cast<float4*>(&s[0]) = ldg4(sn + d0);
          *reinterpret_cast<float4*>(&s[4]) = ldg4(sn + d0 + 4);
          *reinterpret_cast<float4*>(&bl[0]) = ldg4(p.bias + col0 + d0);
          *reinterpret_cast<float4*>(&bl[4]) = ldg4(p.bias + col0 + d0 + 4);
          *reinterpret_cast<float4*>(&bh[0]) = ldg4(p.bias + col0 + 40 + d0);
          *reinterpret_cast<float4*>(&bh[4]) = ldg4(p.bias + col0 + 40 + d0 + 4);
          uint32_t olo[4], ohi[4];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            float x0 = __uint_as_float(lo[j]) + bl[j], x1 = __uint_as_float(lo[j + 1]) + bl[j + 1];
            float y0 = __uint_as_float(hi[j]) + bh[j], y1 = __uint_as_float(hi[j + 1]) + bh[j + 1];
            // rotate_half: out[d] = x*cos - y*sin ; out[d+40] = y*cos + x*sin   (HF :149-167)
            olo[j >> 1] = pack_bf16x2(x0 * c[j] - y0 * s[j], x1 * c[j + 1] - y1 * s[j + 1]);
            ohi[j >> 1] = pack_bf16x2(y0 * c[j] + x0 * s[j], y1 * c[j + 1] + x1 * s[j + 1]);
          }
          *reinterpret_cast<uint4*>(out + d0) = make_uint4(olo[0], olo[1], olo[2], olo[3]);
          *reinterpret_cast<uint4*>(out + 40 + d0) = make_uint4(ohi[0], ohi[1], ohi[2], ohi[3]);
        }
      }
    } else {
#pragma unroll
      for (int d0 = 0; d0 < 80; d0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + d0, v);
        tmem_ld_wait();
        if (row_ok) {
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 b = ldg4(p.bias + col0 + d0 + j);
            o[j >> 1] = pack_bf16x2(__uint_as_float(v[j]) + b.x, __uint_as_float(v[j + 1]) + b.y);
            o[(j >> 1) + 1] = pack_bf16x2(__uint_as_float(v[j + 2]) + b.z, __uint_as_float(v[j + 3]) + b.w);
          }
          *reinterpret_cast<uint4*>(out + d0) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(out + d0 + 8) = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
    }
  } else {
    static_assert(CW % 32 == 0, "generic epilogues work in 32-column chunks");
    const int orow = (row_ok && p.row_map != nullptr) ? p.row_map[row] : row;
#pragma unroll 1
    for (int c = 0; c < CW; c += 32) {
      uint32_t v[32];
      tmem_ld32(taddr + c, v);
      const int col = col0 + c;
      if constexpr (EPI == B200VIT_EPI_BIAS_RESIDUAL) {
        // issue the residual loads before waiting on TMEM so the two latencies overlap
        float4 r[8];
        float* xp = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          r[j] = (row_ok && col + 4 * j + 4 <= p.n) ? *reinterpret_cast<const float4*>(xp + 4 * j)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (col + 4 * j + 4 <= p.n) {
              float4 b = ldg4(p.bias + col + 4 * j);
              r[j].x += __uint_as_float(v[4 * j]) + b.x;
              r[j].y += __uint_as_float(v[4 * j + 1]) + b.y;
              r[j].z += __uint_as_float(v[4 * j + 2]) + b.z;
              r[j].w += __uint_as_float(v[4 * j + 3]) + b.w;
              *reinterpret_cast<float4*>(xp + 4 * j) = r[j];
            }
          }
        }
      } else {
        tmem_ld_wait();
        if (!row_ok) continue;
        if constexpr (EPI == B200VIT_EPI_STORE_F32 || EPI == B200VIT_EPI_BIAS_F32) {
          float* op = reinterpret_cast<float*>(p.out) + static_cast<size_t>(orow) * p.ldo + col;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (col + 4 * j + 4 <= p.n) {
              float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                     __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
              if constexpr (EPI == B200VIT_EPI_BIAS_F32) {
                float4 b = ldg4(p.bias + col + 4 * j);
                o.x += b.x, o.y += b.y, o.z += b.z, o.w += b.w;
              }
              *reinterpret_cast<float4*>(op + 4 * j) = o;
            }
          }
        } else if constexpr (EPI == B200VIT_EPI_SWIGLU) {
          // columns are (gate, up) pairs; 32 accumulator columns -> 16 outputs
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + (col >> 1);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (col + 16 * h + 16 <= p.n) {
              uint32_t o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float4 b = ldg4(p.bias + col + 16 * h + 4 * j);
                const int i = 16 * h + 4 * j;
                float g0 = __uint_as_float(v[i]) + b.x, u0 = __uint_as_float(v[i + 1]) + b.y;
                float g1 = __uint_as_float(v[i + 2]) + b.z, u1 = __uint_as_float(v[i + 3]) + b.w;
                o[j] = pack_bf16x2(silu_f(g0) * u0, silu_f(g1) * u1);
              }
              *reinterpret_cast<uint4*>(op + 8 * h) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
        } else {  // BIAS_GELU, BIAS_BF16
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(orow) * p.ldo + col;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            if (col + 8 * h + 8 <= p.n) {
              float4 b0 = ldg4(p.bias + col + 8 * h), b1 = ldg4(p.bias + col + 8 * h + 4);
              float f[8] = {__uint_as_float(v[8 * h]) + b0.x,     __uint_as_float(v[8 * h + 1]) + b0.y,
                            __uint_as_float(v[8 * h + 2]) + b0.z, __uint_as_float(v[8 * h + 3]) + b0.w,
                            __uint_as_float(v[8 * h + 4]) + b1.x, __uint_as_float(v[8 * h + 5]) + b1.y,
                            __uint_as_float(v[8 * h + 6]) + b1.z, __uint_as_float(v[8 * h + 7]) + b1.w};
              if constexpr (EPI == B200VIT_EPI_BIAS_GELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = gelu_erf_f(f[j]);
              }
              *reinterpret_cast<uint4*>(op + 8 * h) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                                 pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ kernel
template <int BN, int EG, int EPI>
__global__ void __launch_bounds__(128 + 128 * EG, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const GemmParams p) {
  using C = TileCfg<BN>;
  constexpr int STAGES = C::STAGES;
  constexpr int CW = BN / EG;
  static_assert(BN % EG == 0 && BN % 16 == 0 && BN <= 256, "invalid tile");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * C::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_n = (p.n + BN - 1) / BN;
  const int num_m = (p.m + BM - 1) / BM;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.k + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4 * EG);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * BM;
        const int n0 = (tile % num_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], C::STAGE_BYTES);
          tma_load_2d(sA + s * C::A_BYTES, &tma_a, &full[s], kb * BK, m0);
          tma_load_2d(sB + s * C::B_BYTES, &tma_b, &full[s], kb * BK, n0);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty[as], aph ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * C::ACC_STRIDE;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * C::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + s * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_bf16_ss(tacc, umma_desc_k128(a_addr + k * 32), umma_desc_k128(b_addr + k * 32), idesc,
                         (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[s]);  // frees the smem stage when these MMAs retire
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&tfull[as]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;         // TMEM lane quadrant this warp may access
    const int g = (warp - 4) >> 2;  // column group
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m0 = (tile / num_n) * BM;
      const int n0 = (tile % num_n) * BN;
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * C::ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16) + g * CW;
      epilogue_tile<EPI, CW>(taddr, m0 + q * 32 + lane, n0 + g * CW, p);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, int EG, int EPI>
int launch_one(const b200vit_gemm_args& a, cudaStream_t stream) {
  using C = TileCfg<BN>;
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16(&ta, a.d_a, a.m, a.k, BM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tb, a.d_b, a.n, a.k, BN);
  if (rc) return rc;
  GemmParams p{a.d_out, a.d_bias, a.d_row_map, a.d_cos, a.d_sin, a.m, a.n, a.k, a.ldo, a.rope_cols};
  auto kern = gemm_tcgen05_kernel<BN, EG, EPI>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int num_tiles = ((a.m + BM - 1) / BM) * ((a.n + BN - 1) / BN);
  const int sms = device_sm_count();
  const int grid = num_tiles < sms ? num_tiles : sms;
  kern<<<grid, 128 + 128 * EG, C::SMEM_BYTES, stream>>>(ta, tb, p);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

int launch_gemm(const b200vit_gemm_args& a, cudaStream_t stream) {
  if (a.m <= 0 || a.n <= 0 || a.k <= 0) return fail(B200VIT_EINVAL, "gemm: empty problem");
  if (a.k % 8 != 0) return fail(B200VIT_EINVAL, "gemm: K must be a multiple of 8 (16-byte rows for TMA)");
  if ((reinterpret_cast<uintptr_t>(a.d_a) | reinterpret_cast<uintptr_t>(a.d_b) | reinterpret_cast<uintptr_t>(a.d_out)) & 15)
    return fail(B200VIT_EALIGN, "gemm: A, B and out must be 16-byte aligned");
  const bool needs_bias = a.epilogue != B200VIT_EPI_STORE_F32;
  if (needs_bias && a.d_bias == nullptr) return fail(B200VIT_EINVAL, "gemm: epilogue needs a bias vector");
  switch (a.epilogue) {
    case B200VIT_EPI_STORE_F32:
      if (a.n % 8 || a.ldo % 4) return fail(B200VIT_EINVAL, "gemm: N % 8 and ldo % 4 required");
      return launch_one<256, 2, B200VIT_EPI_STORE_F32>(a, stream);
    case B200VIT_EPI_BIAS_F32:
      if (a.n % 8 || a.ldo % 4) return fail(B200VIT_EINVAL, "gemm: N % 8 and ldo % 4 required");
      return launch_one<256, 2, B200VIT_EPI_BIAS_F32>(a, stream);
    case B200VIT_EPI_QKV_ROPE:
      if (a.n % 240 || a.rope_cols % 80 || a.ldo % 8 || !a.d_cos || !a.d_sin)
        return fail(B200VIT_EINVAL, "gemm: QKV epilogue needs N % 240 == 0, head_dim 80, cos/sin tables");
      return launch_one<240, 3, B200VIT_EPI_QKV_ROPE>(a, stream);
    case B200VIT_EPI_BIAS_RESIDUAL:
      if (a.n % 8 || a.ldo % 4) return fail(B200VIT_EINVAL, "gemm: N % 8 and ldo % 4 required");
      return launch_one<256, 2, B200VIT_EPI_BIAS_RESIDUAL>(a, stream);
    case B200VIT_EPI_SWIGLU:
      if (a.n % 16 || a.ldo % 8) return fail(B200VIT_EINVAL, "gemm: SwiGLU needs N % 16 == 0 and ldo % 8 == 0");
      return launch_one<256, 2, B200VIT_EPI_SWIGLU>(a, stream);
    case B200VIT_EPI_BIAS_GELU:
      if (a.n % 8 || a.ldo % 8) return fail(B200VIT_EINVAL, "gemm: N % 8 and ldo % 8 required");
      return launch_one<256, 2, B200VIT_EPI_BIAS_GELU>(a, stream);
    case B200VIT_EPI_BIAS_BF16:
      if (a.n % 8 || a.ldo % 8) return fail(B200VIT_EINVAL, "gemm: N % 8 and ldo % 8 required");
      return launch_one<256, 2, B200VIT_EPI_BIAS_BF16>(a, stream);
    default:
      return fail(B200VIT_EINVAL, "gemm: unknown epilogue");
  }
}

}  // namespace b200
