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
// Structure: persistent CTA pairs (one cluster of 2 per TPC), warp-specialised:
//   warp 0      TMA producer   (own 128 rows of A + half of the B tile, 128-B swizzle, mbarrier ring)
//   warp 1      MMA issuer     (leader CTA, one thread: tcgen05.mma.cta_group::2, 256 x BN x 16,
//                               two TMEM accumulator stages)
//   warp 2      TMEM allocator
//   warps 4..   epilogue       (EG groups of 4 warps: tcgen05.ld -> registers -> swizzled smem ->
//                               TMA store / TMA reduce-add), overlapping the next tile's main loop.
// Global traffic of the hot epilogues goes through the TMA engine: a thread owns one accumulator
// ROW, so direct stores would be 32 separate sectors per instruction (measured: the first version
// of this kernel was epilogue-bound on LSU wavefronts, profiles/r01_*).
//
// RMSNorm (HF :57-71, two per block) has no kernel of its own: RMSNorm(x) W^T == rstd(x) * (x (W diag(gamma))^T).
// gamma is folded into the next weight matrix when the weights are packed (pack.cu); the kernels that WRITE the fp32
// residual stream (patch embed, proj, down: STORE_F32 / BIAS_RESIDUAL_NORM) also write its bf16 copy -- the next
// GEMM's A operand -- and per-row partial sums of x^2; the kernels that READ it (QKV_ROPE, SWIGLU) add the partials
// in index order and scale their accumulator rows by rstd before the bias.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "internal.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace b200 {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int SMEM_LIMIT = 227 * 1024;

struct GemmParams {
  void* out;
  const float* bias;
  const int32_t* row_map;
  const float2* rope;     // [P, 20] fp32 (cos, sin) of coordinate * inv_freq, P = largest grid side
  const int2* rope_pos;   // [M] (hpos, wpos) of every row
  int m, n, k, ldo;
  int stream_k;  // bit 0: stream-K decomposition (BIAS_RESIDUAL_NORM only); bit 1: no weight prefetch before the PDL wait;
                 // bit 2: QKV_ROPE_WINATTN does NOT hold the tcgen05 issue back while its warp-level MMAs run
  // fused RMSNorm (HF :57-71): RMSNorm(x) W^T == rstd(x) * (x (W diag(gamma))^T).  Producers of the fp32 residual
  // stream also emit its bf16 copy (the next GEMM's A operand) and per-row partial sums of x^2, one per 128-column
  // group, laid out [part][M]; consumers add the partials in index order and scale their accumulator rows by rstd.
  __nv_bfloat16* out_bf16;
  float* rowsq_out;
  const float* rowsq_in;
  int rowsq_parts;
  float norm_eps;
  int32_t* sync;  // stream-K hand-over flags, one per (unit, CTA rank, epilogue warp); zero between launches
};

// sum of the row's partials in index order (bit-stable) -> rsqrt(mean(x^2) + eps)
// All loads are issued before the first add (a rolled loop would serialise up to 16 L2 round trips in front of
// every tile's epilogue); callers ask for it BEFORE they wait for the accumulator.
constexpr int MAX_ROWSQ_PARTS = 16;
__device__ __forceinline__ float row_rstd(const GemmParams& p, int row) {
  if (p.rowsq_in == nullptr) return 1.0f;
  float v[MAX_ROWSQ_PARTS];
#pragma unroll
  for (int i = 0; i < MAX_ROWSQ_PARTS; ++i)
    v[i] = i < p.rowsq_parts ? __ldg(p.rowsq_in + static_cast<size_t>(i) * p.m + row) : 0.f;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_ROWSQ_PARTS; ++i) s += v[i];
  return rsqrtf(s / static_cast<float>(p.k) + p.norm_eps);
}

__host__ __device__ constexpr bool is_resid_norm(int epi) { return epi == B200VIT_EPI_BIAS_RESIDUAL_NORM; }

#ifndef B200_RESID_BUFS
#define B200_RESID_BUFS 1
#endif
constexpr int RESID_BUFS = B200_RESID_BUFS;  // staging buffers per warp of the reduce-add epilogue
// staging bytes per epilogue warp (TMA-store epilogues), 0 = direct global stores
__host__ __device__ constexpr int epi_stage_bytes(int epi) {
  return epi == B200VIT_EPI_QKV_ROPE_WINATTN ? 32 * 176  // padded rows: conflict-free ldmatrix of the staged Q, K, V
         : epi == B200VIT_EPI_QKV_ROPE      ? 32 * 160
         : epi == B200VIT_EPI_BIAS_RESIDUAL ? RESID_BUFS * 4096
         : is_resid_norm(epi) ? 4096  // one 32 x 32 fp32 chunk of (acc + bias) in transit between the two thread layouts
         : epi == B200VIT_EPI_SWIGLU        ? 4096
         : epi == B200VIT_EPI_BIAS_GELU     ? 2 * 4096
                                            : 4096;  // row-mapped epilogues: one chunk in transit between thread layouts
}

// PAIR = true: two CTAs of a cluster cooperate on a 256 x BN tile with tcgen05.mma.cta_group::2 --
// each CTA stages its own 128 rows of A and HALF of the B tile (1.5x less L2->SM operand traffic).
template <int BN, int EG, int EPI, bool PAIR>
struct TileCfg {
  static constexpr int BN_LOAD = PAIR ? BN / 2 : BN;  // B rows staged by one CTA
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN_LOAD * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_WARPS = 4 * EG;
  // swizzled TMA staging wants 1024-byte alignment; the fused-attention staging is only read by ldmatrix / a dense
  // TMA box (128-byte alignment) and keeps its exact size so that five operand stages still fit
  static constexpr int STG_WARP = EPI == B200VIT_EPI_QKV_ROPE_WINATTN ? epi_stage_bytes(EPI) : (epi_stage_bytes(EPI) + 1023) / 1024 * 1024;
  static constexpr int STG_BYTES = EPI_WARPS * STG_WARP;
  static constexpr int BAR_BYTES = 256;
  static constexpr int AVAIL = SMEM_LIMIT - 1024 - BAR_BYTES - STG_BYTES;
  static constexpr int STAGES = AVAIL / STAGE_BYTES > 8 ? 8 : AVAIL / STAGE_BYTES;
  static constexpr int ACC_STRIDE = (BN <= 128) ? 128 : 256;  // TMEM columns between the two accumulators
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + BAR_BYTES + 1024;
  static_assert(STAGES >= 3, "not enough shared memory for the pipeline");
  static_assert((2 * STAGES + 4) * 8 + 16 <= BAR_BYTES, "barrier area too small");
};

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 bias4(const float* bias, int col, int n) {
  return (col + 4 <= n) ? ldg4(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// One row of one head of Q, K or V: rstd * acc + bias, 2-D RoPE for Q and K (HF :149-167), bf16, 160 bytes at `srow`.
// `rs` = fused RMSNorm: per-row rstd of the A operand's source.  `bcol` = first of the head's 80 bias entries.
__device__ __forceinline__ void qkv_row_to_smem(uint32_t taddr, int row, bool row_ok, int bcol, bool rope, const GemmParams& p,
                                                float rs, uint32_t srow) {
  if (rope) {
    // HF :382-409: rotary_pos_emb = freqs[pos_ids] with freqs = outer(arange(max_grid), inv_freq): dims 0..19 turn
    // with the row's hpos, dims 20..39 with its wpos.  The [P, 20] table stays in L1; only 8 bytes per row are new.
    const int2 pos = row_ok ? __ldg(p.rope_pos + row) : make_int2(0, 0);
    const float4* th = reinterpret_cast<const float4*>(p.rope + static_cast<size_t>(pos.x) * 20);
    const float4* tv = reinterpret_cast<const float4*>(p.rope + static_cast<size_t>(pos.y) * 20);
#pragma unroll
    for (int d0 = 0; d0 < 40; d0 += 8) {
      uint32_t lo[8], hi[8];
      tmem_ld8(taddr + d0, lo);
      tmem_ld8(taddr + 40 + d0, hi);
      float4 tw[4];  // (cos, sin) pairs of dims d0 .. d0+7, fp32 as HF rotates (:149-167)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = d0 + 2 * i;
        tw[i] = d < 20 ? __ldg(th + (d >> 1)) : __ldg(tv + ((d - 20) >> 1));
      }
      float bl[8], bh[8];
      *reinterpret_cast<float4*>(&bl[0]) = ldg4(p.bias + bcol + d0);
      *reinterpret_cast<float4*>(&bl[4]) = ldg4(p.bias + bcol + d0 + 4);
      *reinterpret_cast<float4*>(&bh[0]) = ldg4(p.bias + bcol + 40 + d0);
      *reinterpret_cast<float4*>(&bh[4]) = ldg4(p.bias + bcol + 40 + d0 + 4);
      tmem_ld_wait();
      uint32_t olo[4], ohi[4];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float c0 = tw[j >> 1].x, s0 = tw[j >> 1].y, c1 = tw[j >> 1].z, s1 = tw[j >> 1].w;
        const float x0 = __uint_as_float(lo[j]) * rs + bl[j], x1 = __uint_as_float(lo[j + 1]) * rs + bl[j + 1];
        const float y0 = __uint_as_float(hi[j]) * rs + bh[j], y1 = __uint_as_float(hi[j + 1]) * rs + bh[j + 1];
        // rotate_half: out[d] = x*cos - y*sin ; out[d+40] = y*cos + x*sin
        olo[j >> 1] = pack_bf16x2(x0 * c0 - y0 * s0, x1 * c1 - y1 * s1);
        ohi[j >> 1] = pack_bf16x2(y0 * c0 + x0 * s0, y1 * c1 + x1 * s1);
      }
      st_shared_v4(srow + d0 * 2, olo[0], olo[1], olo[2], olo[3]);
      st_shared_v4(srow + 80 + d0 * 2, ohi[0], ohi[1], ohi[2], ohi[3]);
    }
  } else {
#pragma unroll
    for (int d0 = 0; d0 < 80; d0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + d0, v);
      tmem_ld_wait();
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b = ldg4(p.bias + bcol + d0 + j);
        o[j >> 1] = pack_bf16x2(__uint_as_float(v[j]) * rs + b.x, __uint_as_float(v[j + 1]) * rs + b.y);
        o[(j >> 1) + 1] = pack_bf16x2(__uint_as_float(v[j + 2]) * rs + b.z, __uint_as_float(v[j + 3]) * rs + b.w);
      }
      st_shared_v4(srow + d0 * 2, o[0], o[1], o[2], o[3]);
      st_shared_v4(srow + d0 * 2 + 16, o[4], o[5], o[6], o[7]);
    }
  }
}

// ------------------------------------------------------------------ epilogues
// Each epilogue thread owns one accumulator row (TMEM lane) and CW consecutive columns.
// `stg` = this warp's staging buffer (shared address), `row0` = first row of the warp's 32-row slab.
template <int EPI, int CW>
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, int row0, int lane, int col0, const GemmParams& p,
                                              const CUtensorMap* tma_out, uint32_t stg, bool first_split, float rs) {
  const int row = row0 + lane;
  const bool row_ok = row < p.m;
  if constexpr (EPI == B200VIT_EPI_QKV_ROPE) {
    static_assert(CW == 80, "QKV epilogue handles one 80-wide head per warp group");
    if (lane == 0) bulk_wait_read<0>();  // previous tile's store has drained the staging buffer
    __syncwarp();
    // weight rows are head-interleaved [(head, {q, k, v}, 80)]: tile column col0 = head * 240 + type * 80
    const int head = col0 / 240, type = (col0 % 240) / 80;
    qkv_row_to_smem(taddr, row_ok ? row : 0, row_ok, col0, type < 2, p, rs, stg + lane * 160);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tma_out, stg, type * (p.n / 3) + head * 80, row0);  // the qkv buffer stays Q | K | V, head-major
      bulk_commit();
    }
  } else if constexpr (EPI == B200VIT_EPI_BIAS_RESIDUAL) {
    // x += acc + bias through TMA reduce-add (fp32 add performed at L2): no residual loads on the SM.
    static_assert(CW % 32 == 0, "32-column chunks");
#pragma unroll
    for (int c = 0; c < CW; c += 32) {
      const uint32_t buf = stg + ((c >> 5) % RESID_BUFS) * 4096;
      uint32_t v[32];
      tmem_ld32(taddr + c, v);
      if (lane == 0) bulk_wait_read<RESID_BUFS - 1>();  // the store that last used this buffer has read it
      __syncwarp();
      const int col = col0 + c;
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // a tile completed by several stream-K partials gets the bias from the split that holds k = 0
        const float4 b = first_split ? bias4(p.bias, col + 4 * j, p.n) : make_float4(0.f, 0.f, 0.f, 0.f);
        st_shared_v4(swz128(buf, lane, j), __float_as_uint(__uint_as_float(v[4 * j]) + b.x),
                     __float_as_uint(__uint_as_float(v[4 * j + 1]) + b.y),
                     __float_as_uint(__uint_as_float(v[4 * j + 2]) + b.z),
                     __float_as_uint(__uint_as_float(v[4 * j + 3]) + b.w));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_2d(tma_out, buf, col, row0);
        bulk_commit();
      }
    }
  } else if constexpr (EPI == B200VIT_EPI_SWIGLU) {
    // columns are (gate, up) pairs; CW = 128 accumulator columns -> 64 bf16 outputs = one 128-byte box row
    static_assert(CW == 128, "SwiGLU epilogue expects 128 accumulator columns per warp");
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
#pragma unroll
    for (int c = 0; c < CW; c += 32) {
      uint32_t v[32];
      tmem_ld32(taddr + c, v);
      const int col = col0 + c;
      float4 b[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = bias4(p.bias, col + 4 * j, p.n);
      tmem_ld_wait();
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float g0 = __uint_as_float(v[4 * j]) * rs + b[j].x, u0 = __uint_as_float(v[4 * j + 1]) * rs + b[j].y;
        const float g1 = __uint_as_float(v[4 * j + 2]) * rs + b[j].z, u1 = __uint_as_float(v[4 * j + 3]) * rs + b[j].w;
        o[j] = pack_bf16x2(silu_f(g0) * u0, silu_f(g1) * u1);
      }
      st_shared_v4(swz128(stg, lane, (c >> 4)), o[0], o[1], o[2], o[3]);
      st_shared_v4(swz128(stg, lane, (c >> 4) + 1), o[4], o[5], o[6], o[7]);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tma_out, stg, col0 >> 1, row0);
      bulk_commit();
    }
  } else if constexpr (EPI == B200VIT_EPI_BIAS_GELU) {
    static_assert(CW % 64 == 0, "64-column boxes");
#pragma unroll
    for (int c = 0; c < CW; c += 64) {
      const uint32_t buf = stg + ((c >> 6) & 1) * 4096;
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld32(taddr + c + 32 * h, v);
        const int col = col0 + c + 32 * h;
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b0 = bias4(p.bias, col + 8 * j, p.n), b1 = bias4(p.bias, col + 8 * j + 4, p.n);
          st_shared_v4(swz128(buf, lane, 4 * h + j),
                       pack_bf16x2(gelu_erf_f(__uint_as_float(v[8 * j]) + b0.x), gelu_erf_f(__uint_as_float(v[8 * j + 1]) + b0.y)),
                       pack_bf16x2(gelu_erf_f(__uint_as_float(v[8 * j + 2]) + b0.z), gelu_erf_f(__uint_as_float(v[8 * j + 3]) + b0.w)),
                       pack_bf16x2(gelu_erf_f(__uint_as_float(v[8 * j + 4]) + b1.x), gelu_erf_f(__uint_as_float(v[8 * j + 5]) + b1.y)),
                       pack_bf16x2(gelu_erf_f(__uint_as_float(v[8 * j + 6]) + b1.z), gelu_erf_f(__uint_as_float(v[8 * j + 7]) + b1.w)));
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tma_out, buf, col0 + c, row0);
        bulk_commit();
      }
    }
  } else {
    // Row-mapped outputs (window reorder of the patch embed, un-reorder of the merger).  A thread owns one accumulator
    // ROW, but a warp store should cover whole 128-byte lines: every 32 x 32 chunk is transposed through 4 KB of
    // swizzled shared memory (written row-per-thread, read back 8 lanes per row), then each row's 128 bytes leave in
    // one piece to wherever row_map sends that row.  STORE_F32 also emits the bf16 copy and the row's sum of squares
    // (the first fused RMSNorm's inputs).
    static_assert(CW % 32 == 0, "generic epilogues work in 32-column chunks");
    const int lr = lane >> 3, lc = (lane & 7) * 4;  // coalesced layout: row 4 i + lr (i = 0..7), columns lc .. lc + 3
    int orow[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = row0 + 4 * i + lr;
      orow[i] = r < p.m ? (p.row_map != nullptr ? __ldg(p.row_map + r) : r) : -1;
    }
    float ss[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ss[i] = 0.f;
#pragma unroll 1
    for (int c = 0; c < CW; c += 32) {
      uint32_t v[32];
      tmem_ld32(taddr + c, v);
      const int col = col0 + c + lc;
      const bool col_ok = col + 4 <= p.n;
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (EPI != B200VIT_EPI_STORE_F32) bb = bias4(p.bias, col, p.n);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) st_shared_v4(swz128(stg, lane, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = ld_shared_f4(swz128(stg, 4 * i + lr, lane & 7));
        const float4 o = make_float4(a.x + bb.x, a.y + bb.y, a.z + bb.z, a.w + bb.w);
        if (orow[i] < 0 || !col_ok) continue;
        const size_t off = static_cast<size_t>(orow[i]) * p.ldo + col;
        if constexpr (EPI == B200VIT_EPI_BIAS_BF16) {
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        } else {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = o;
          if constexpr (EPI == B200VIT_EPI_STORE_F32) {
            ss[i] += o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
            if (p.out_bf16 != nullptr)
              *reinterpret_cast<uint2*>(p.out_bf16 + off) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
          }
        }
      }
      __syncwarp();  // every lane has read the chunk back before the next one overwrites it
    }
    if constexpr (EPI == B200VIT_EPI_STORE_F32) {
      // row sums: the 8 lanes that share a row combine their column partials in a fixed tree
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float s8 = ss[i];
        s8 += __shfl_xor_sync(0xffffffffu, s8, 1);
        s8 += __shfl_xor_sync(0xffffffffu, s8, 2);
        s8 += __shfl_xor_sync(0xffffffffu, s8, 4);
        if ((lane & 7) == 0 && p.rowsq_out != nullptr && orow[i] >= 0 && col0 < p.n)
          p.rowsq_out[static_cast<size_t>(col0 / CW) * p.m + orow[i]] = s8;
      }
    }
  }
}

// ------------------------------------------------------------------ work decomposition
// A "segment" is a contiguous range of k-blocks of one output tile.  Data-parallel mode: unit u
// takes whole tiles u, u + U, ...  Stream-K mode (BIAS_RESIDUAL_NORM only): the flattened
// (tile, k-block) space is cut into U equal ranges, so all CTA pairs finish together even when
// the tile count is a poor multiple of the pair count (N = 1280 at M = 8192: 160 tiles on 74
// pairs); a tile cut by a range boundary is finished by exactly two units in a fixed order (tail
// partial reduce-added first, then the head unit's epilogue; see the epilogue for the hand-over).
struct Segment {
  int tile, kb0, kb1;
};
struct WorkIter {
  int num_tiles, num_kb, unit, num_units;
  bool stream_k;
  int tile;               // data-parallel cursor
  long long pos, end;     // stream-K cursor in k-block units
  __device__ WorkIter(int num_tiles_, int num_kb_, int unit_, int num_units_, bool stream_k_)
      : num_tiles(num_tiles_), num_kb(num_kb_), unit(unit_), num_units(num_units_), stream_k(stream_k_), tile(unit_) {
    if (stream_k) {
      pos = boundary(unit);
      end = boundary(unit + 1);
    }
  }
  // range boundary of unit u, snapped to a tile boundary when it would leave a sliver (< 4 k-blocks)
  __device__ long long boundary(int u) const {
    const long long total = static_cast<long long>(num_tiles) * num_kb;
    if (u >= num_units) return total;
    long long b = total * u / num_units;
    const int r = static_cast<int>(b % num_kb);
    if (r < 4) b -= r;
    else if (r > num_kb - 4) b += num_kb - r;
    return b;
  }
  __device__ bool next(Segment& s) {
    if (!stream_k) {
      if (tile >= num_tiles) return false;
      s.tile = tile, s.kb0 = 0, s.kb1 = num_kb;
      tile += num_units;
      return true;
    }
    if (pos >= end) return false;
    s.tile = static_cast<int>(pos / num_kb);
    s.kb0 = static_cast<int>(pos % num_kb);
    const long long tile_end = static_cast<long long>(s.tile + 1) * num_kb;
    const long long e = end < tile_end ? end : tile_end;
    s.kb1 = s.kb0 + static_cast<int>(e - pos);
    pos = e;
    return true;
  }
};

// ------------------------------------------------------------------ kernel
template <int BN, int EG, int EPI, bool PAIR>
__global__ void __launch_bounds__(128 + 128 * EG, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_aux,
                    const GemmParams p) {
  using C = TileCfg<BN, EG, EPI, PAIR>;
  constexpr int STAGES = C::STAGES;
  constexpr int CW = BN / EG;
  constexpr int NCTA = PAIR ? 2 : 1;
  constexpr int MT = BM * NCTA;  // rows of C per scheduled tile
  static_assert(BN % EG == 0 && BN % 16 == 0 && BN <= 256, "invalid tile");
  static_assert(C::B_BYTES % 1024 == 0, "B tile must keep 1024-byte stage alignment");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * C::A_BYTES;
  uint8_t* sStg = smem + STAGES * C::STAGE_BYTES;  // 1024-aligned (stage sizes are multiples of 1024)
  uint64_t* full = reinterpret_cast<uint64_t*>(sStg + C::STG_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  // QKV_ROPE_WINATTN: raised by the leader CTA's epilogue while its warp-level MMAs run; the MMA thread holds the tcgen05
  // issue back meanwhile (see the epilogue)
  volatile uint32_t* hmma_gate = tmem_slot + 1;

  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs)
  const int unit = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int num_units = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const int num_n = (p.n + BN - 1) / BN;
  const int num_m = (p.m + MT - 1) / MT;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.k + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (epi_stage_bytes(EPI) > 0) tma_prefetch_desc(&tma_out);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], NCTA);  // one arrival per producing CTA (+ the TMA transaction bytes)
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4 * EG * NCTA);  // one arrival per epilogue warp of every CTA
    }
    *hmma_gate = 0;
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, C::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above overlapped the previous kernel's tail.  Each role executes
  // griddepcontrol.wait itself before it touches memory the previous kernel produced (A operand, output buffer);
  // the producer first prefetches the WEIGHT tiles of its first pipeline stages, which depend on nothing.

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer (every CTA) =====================
#ifdef B200_GEMM_DBG_NOTMA
      if (true) { griddep_wait(); } else  // measurement only: no loads at all, the MMAs run on whatever is in shared memory
#endif
      {
      // phase 1 (before the dependency wait): arm the first stages and start their B (weight) loads
      int pre = 0;
      {
        WorkIter w0(num_tiles, num_kb, unit, num_units, (p.stream_k & 1) != 0);
        Segment s0;
        while (!(p.stream_k & 2) && pre < STAGES && w0.next(s0)) {
          const int n0 = (s0.tile % num_n) * BN + rank * C::BN_LOAD;
          for (int kb = s0.kb0; kb < s0.kb1 && pre < STAGES; ++kb, ++pre) {
            if constexpr (PAIR) {
              const uint32_t lead_full = mapa_u32(smem_u32(&full[pre]), 0);
              if (rank == 0) mbar_arrive_expect_tx(&full[pre], 2 * C::STAGE_BYTES);
              tma_load_2d_pair(sB + pre * C::B_BYTES, &tma_b, lead_full, kb * BK, n0);
            } else {
              mbar_arrive_expect_tx(&full[pre], C::STAGE_BYTES);
              tma_load_2d(sB + pre * C::B_BYTES, &tma_b, &full[pre], kb * BK, n0);
            }
          }
        }
      }
      griddep_wait();
      // phase 2: the regular ring; the first `pre` k-blocks only need their A tile
      int s = 0;
      uint32_t ph = 0;
      int issued = 0;
      WorkIter work(num_tiles, num_kb, unit, num_units, (p.stream_k & 1) != 0);
      Segment sg;
      while (work.next(sg)) {
        const int m0 = (sg.tile / num_n) * MT + rank * BM;
        const int n0 = (sg.tile % num_n) * BN + rank * C::BN_LOAD;
        for (int kb = sg.kb0; kb < sg.kb1; ++kb, ++issued) {
          const bool b_done = issued < pre;
          if (!b_done) mbar_wait(&empty[s], ph ^ 1);
          if constexpr (PAIR) {
            const uint32_t lead_full = mapa_u32(smem_u32(&full[s]), 0);
            if (!b_done && rank == 0) mbar_arrive_expect_tx(&full[s], 2 * C::STAGE_BYTES);
            tma_load_2d_pair(sA + s * C::A_BYTES, &tma_a, lead_full, kb * BK, m0);
            if (!b_done) tma_load_2d_pair(sB + s * C::B_BYTES, &tma_b, lead_full, kb * BK, n0);
            if (rank != 0) mbar_arrive_cluster(lead_full);
          } else {
            if (!b_done) mbar_arrive_expect_tx(&full[s], C::STAGE_BYTES);
            tma_load_2d(sA + s * C::A_BYTES, &tma_a, &full[s], kb * BK, m0);
            if (!b_done) tma_load_2d(sB + s * C::B_BYTES, &tma_b, &full[s], kb * BK, n0);
          }
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===================== MMA issuer (leader CTA) =====================
      constexpr uint32_t idesc = umma_idesc_bf16(MT, BN);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      WorkIter work(num_tiles, num_kb, unit, num_units, (p.stream_k & 1) != 0);
      Segment sg;
#ifdef B200_GEMM_TIMING
      long long t_acc = 0, t_full = 0, t_issue = 0, t_gate = 0, t_prev = clock64(), t_start = t_prev;
      int n_seg = 0, n_kb = 0, n_held = 0;
#define GSTAMP(v) do { long long _n = clock64(); v += _n - t_prev; t_prev = _n; } while (0)
#else
#define GSTAMP(v)
#endif
      for (; work.next(sg); ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        GSTAMP(t_issue);
        mbar_wait(&tempty[as], aph ^ 1);  // epilogues have drained this accumulator
        tc_fence_after();
        GSTAMP(t_acc);
#ifdef B200_GEMM_TIMING
        ++n_seg; n_kb += sg.kb1 - sg.kb0;
#endif
        const uint32_t tacc = tmem_base + as * C::ACC_STRIDE;
        for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
          GSTAMP(t_issue);
          // the gate is sampled BEFORE the stage wait: a shared-memory load takes ~200 cycles while the operand traffic
          // saturates the port, and behind the wait it would sit on the issue path of every k-block
          uint32_t gate_up = 0;
          if constexpr (EPI == B200VIT_EPI_QKV_ROPE_WINATTN) gate_up = *hmma_gate;
#ifndef B200_GEMM_DBG_NOTMA
          mbar_wait(&full[s], ph);
          tc_fence_after();
#endif
          GSTAMP(t_full);
          if constexpr (EPI == B200VIT_EPI_QKV_ROPE_WINATTN) {
            if (gate_up != 0) {
#ifdef B200_GEMM_TIMING
              ++n_held;
#endif
              while (*hmma_gate != 0) {  // the epilogue's mma.sync section has the tensor cores to itself
              }
            }
            GSTAMP(t_gate);
          }
          const uint32_t a_addr = smem_u32(sA + s * C::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + s * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if constexpr (PAIR)
              umma_bf16_ss_pair(tacc, umma_desc_k128(a_addr + k * 32), umma_desc_k128(b_addr + k * 32), idesc,
                                (kb != sg.kb0 || k != 0) ? 1u : 0u);
            else
              umma_bf16_ss(tacc, umma_desc_k128(a_addr + k * 32), umma_desc_k128(b_addr + k * 32), idesc,
                           (kb != sg.kb0 || k != 0) ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs) when these MMAs retire
          if constexpr (PAIR) umma_commit_pair(&empty[s], 0x3); else umma_commit(&empty[s]);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        // accumulator complete -> epilogues of both CTAs
        if constexpr (PAIR) umma_commit_pair(&tfull[as], 0x3); else umma_commit(&tfull[as]);
      }
#ifdef B200_GEMM_TIMING
      GSTAMP(t_issue);
      if (blockIdx.x == 10) printf("gemm mma thread: %d segs %d kb | total %lld | wait tempty %lld | wait full %lld | issue %lld | held by the epilogue %lld (%d times)\n",
                                   n_seg, n_kb, clock64() - t_start, t_acc, t_full, t_issue, t_gate, n_held);
#endif
    }
  } else if (warp >= 4) {
    // ===================== epilogue (every CTA, its own 128 rows) =====================
    griddep_wait();                 // the output buffer may still be in use by the previous kernel
    const int q = warp & 3;         // TMEM lane quadrant this warp may access
    const int g = (warp - 4) >> 2;  // column group
    const uint32_t stg = smem_u32(sStg) + (warp - 4) * C::STG_WARP;
    int it = 0;
    WorkIter work(num_tiles, num_kb, unit, num_units, (p.stream_k & 1) != 0);
    Segment sg;
#ifdef B200_GEMM_TIMING
    long long e_wait = 0, e_busy = 0, e_a = 0, e_b = 0, e_c = 0, e_prev = clock64(), e_start = e_prev, e_lastfull = 0;
#define ESTAMP(v) do { long long _n = clock64(); v += _n - e_prev; e_prev = _n; } while (0)
#else
#define ESTAMP(v)
#endif
    if constexpr (EPI == B200VIT_EPI_QKV_ROPE_WINATTN) {
      // QKV projection + RoPE + the WINDOW attention of the 28 windowed layers in one kernel.  The weight rows are
      // head-interleaved, so the 240 columns of a tile are Q_h | K_h | V_h of one head h, and a CTA's 128 rows are two
      // 64-patch windows: everything softmax(Q K^T / sqrt(80)) V needs for (2 windows x 1 head) is in this CTA's
      // accumulator.  The 12 epilogue warps stage Q, K (rotated) and V as bf16 in shared memory exactly as the plain
      // epilogue does -- and then, instead of storing them, eight warps run the attention of 16 query rows each with
      // warp-level MMAs (5.2 MFLOP per tile against 78.6 MFLOP of projection) and store the output rows.
      // Q, K, V never reach L2 / HBM (63 MB written and read back per layer otherwise) and the 28 attention launches
      // disappear.  Overlaps the next tile's main loop like every other epilogue.
      static_assert(CW == 80 && EG == 3, "one head per tile: Q | K | V groups of 80 columns");
      constexpr int RS = 176;  // staged row stride in bytes
      constexpr int EPI_THREADS = 128 * EG;
      const uint32_t stg_all = smem_u32(sStg);
      const float scale_log2 = 0.11180339887498948f * 1.4426950408889634f;  // 80^-0.5 * log2(e)
      for (; work.next(sg); ++it) {
        const int m0 = (sg.tile / num_n) * MT + rank * BM;
        const int head = sg.tile % num_n;
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        const int row = m0 + q * 32 + lane;
        const float rs = row_rstd(p, row < p.m ? row : 0);
        ESTAMP(e_busy);
        mbar_wait(&tfull[as], aph);
        tc_fence_after();
        ESTAMP(e_wait);
        const uint32_t taddr = tmem_base + as * C::ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16) + g * CW;
        if (lane == 0) bulk_wait_read<0>();  // the previous tile's output store has drained this warp's buffer
        __syncwarp();
        qkv_row_to_smem(taddr, row < p.m ? row : 0, row < p.m, head * 240 + g * 80, g < 2, p, rs, stg + lane * RS);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[as]), 0));
          else mbar_arrive(&tempty[as]);
        }
        ESTAMP(e_a);
        named_barrier(1, EPI_THREADS);  // Q, K, V of the CTA's two windows are staged
        // mma.sync and tcgen05.mma share the tensor cores, and while the main loop runs the warp-level MMAs get ~1/8 of
        // their stand-alone rate (70 cycles per HMMA and sub-partition, whether four warps or eight issue them): 11 k
        // cycles of attention per tile against a 10 k cycle main loop that cannot hide them.  So the leader CTA's
        // epilogue raises a gate that holds the tcgen05 issue back for the few thousand cycles the section then needs
        // (the partner CTA's section runs at the same time: both are released by the same multicast commit).
        const bool gate_owner = rank == 0 && warp == 4 && lane == 0 && !(p.stream_k & 4);
        if (gate_owner) *hmma_gate = 1;
#ifdef B200_GEMM_TIMING
        const long long gate_t0 = clock64();
#endif
        ESTAMP(e_b);
        // The epilogue, not the main loop, is what this kernel waits for (MMA-thread clocks: 33 k of 116 k cycles blocked on
        // tempty when four warps ran the attention, 11 k cycles per tile at ~70 cycles per dependent ldmatrix -> HMMA step
        // under the main loop's saturated shared-memory port).  So EIGHT warps run it, 16 query rows each: warp (g, q),
        // g < 2, takes rows 16 g .. 16 g + 15 of quadrant q's 32 staged Q rows.  K / V fragments are read once per warp
        // (four times per window instead of twice): +80 KB of ldmatrix traffic per tile against 1.6 MB of operand traffic.
        uint32_t opk[20];              // this warp's 16 x 80 output, bf16 pairs: [dim tile][row half]
        const uint32_t qbuf = stg_all + q * C::STG_WARP;  // quadrant q's Q rows (staged by warp (0, q)); later its output tile
        if (g < 2) {
          const int w = q >> 1;         // window (rows 64 w .. 64 w + 63 of the CTA) of this warp's query rows
          const uint32_t qb = qbuf + 16 * g * RS;
          const int lr = (lane & 7) + 8 * ((lane >> 3) & 1), lc = lane >> 4;     // ldmatrix address roles
          float sacc[8][4];
#pragma unroll
          for (int i = 0; i < 8; ++i) sacc[i][0] = sacc[i][1] = sacc[i][2] = sacc[i][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) {  // S = Q K^T over the 80 head dims, 16 at a time
            uint32_t a[4];
            ldmatrix_x4(qb + lr * RS + (ks * 16 + 8 * lc) * 2, a[0], a[1], a[2], a[3]);
#pragma unroll
            for (int nt = 0; nt < 8; nt += 2) {  // keys 8 nt .. 8 nt + 15
              const uint32_t kb = stg_all + (1 * 4 + 2 * w + (nt >> 2)) * C::STG_WARP + ((8 * nt) & 31) * RS;
              uint32_t b0, b1, b2, b3;
              ldmatrix_x4(kb + ((lane & 7) + 8 * lc) * RS + (ks * 16 + 8 * ((lane >> 3) & 1)) * 2, b0, b1, b2, b3);
              mma_bf16_16816(sacc[nt], a[0], a[1], a[2], a[3], b0, b1);
              mma_bf16_16816(sacc[nt + 1], a[0], a[1], a[2], a[3], b2, b3);
            }
          }
          // softmax over the window's 64 keys; a thread holds rows lane/4 and lane/4 + 8, a quad a whole row
          uint32_t pa[4][4];  // P as A operands of the four 16-key steps
          float inv[2];
          {
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              mx0 = fmaxf(mx0, fmaxf(sacc[i][0], sacc[i][1]));
              mx1 = fmaxf(mx1, fmaxf(sacc[i][2], sacc[i][3]));
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)), mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)), mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float e0 = ex2f((sacc[i][0] - mx0) * scale_log2), e1 = ex2f((sacc[i][1] - mx0) * scale_log2);
              const float e2 = ex2f((sacc[i][2] - mx1) * scale_log2), e3 = ex2f((sacc[i][3] - mx1) * scale_log2);
              sum0 += e0 + e1, sum1 += e2 + e3;
              pa[i >> 1][(i & 1) * 2] = pack_bf16x2(e0, e1);
              pa[i >> 1][(i & 1) * 2 + 1] = pack_bf16x2(e2, e3);
            }
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1), sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1), sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
            inv[0] = 1.0f / sum0, inv[1] = 1.0f / sum1;
          }
#pragma unroll
          for (int dp = 0; dp < 5; ++dp) {  // O = P V, two 8-wide dim tiles per pass
            float oacc[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // keys 16 j .. 16 j + 15
              const uint32_t vb = stg_all + (2 * 4 + 2 * w + (j >> 1)) * C::STG_WARP + ((16 * j) & 31) * RS;
              uint32_t b0, b1, b2, b3;
              ldmatrix_x4_trans(vb + lr * RS + (dp * 16 + 8 * lc) * 2, b0, b1, b2, b3);
              mma_bf16_16816(oacc[0], pa[j][0], pa[j][1], pa[j][2], pa[j][3], b0, b1);
              mma_bf16_16816(oacc[1], pa[j][0], pa[j][1], pa[j][2], pa[j][3], b2, b3);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              opk[(dp * 2 + i) * 2] = pack_bf16x2(oacc[i][0] * inv[0], oacc[i][1] * inv[0]);
              opk[(dp * 2 + i) * 2 + 1] = pack_bf16x2(oacc[i][2] * inv[1], oacc[i][3] * inv[1]);
            }
          }
        }
        ESTAMP(e_c);
        named_barrier(2, EPI_THREADS);  // every read of the staged Q, K, V is done: the buffers may be reused
        if (gate_owner) *hmma_gate = 0;
#ifdef B200_GEMM_TIMING
        if (gate_owner && blockIdx.x == 10) printf("gate up %lld cycles\n", clock64() - gate_t0);
#endif
        ESTAMP(e_b);
        if (g < 2) {
          // 32 x 80 bf16 per quadrant, dense 160-byte rows at the start of the Q warp's buffer (rows 16 g .. from this
          // warp) -> one TMA store by the Q warp once both halves are in
#pragma unroll
          for (int i = 0; i < 10; ++i) {
            const uint32_t base = qbuf + (16 * g + (lane >> 2)) * 160 + (i * 8 + 2 * (lane & 3)) * 2;
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(base), "r"(opk[2 * i]) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 8 * 160), "r"(opk[2 * i + 1]) : "memory");
          }
          fence_proxy_async_smem();
          named_barrier(3 + q, 64);
          if (g == 0 && lane == 0) {
            tma_store_2d(&tma_out, qbuf, head * 80, m0 + q * 32);
            bulk_commit();
          }
        }
      }
    } else
    if constexpr (is_resid_norm(EPI)) {
      // x += acc + bias with the NEW x in hand: besides the fp32 residual stream the epilogue writes its bf16 copy
      // (A operand of the next GEMM) and per-row sums of squares (the next RMSNorm's statistics).
      // A thread can only read ITS row of the accumulator from TMEM, but global memory wants the opposite layout
      // (a warp instruction should cover whole 128-byte lines), so each 32 x 32 chunk of acc + bias is transposed
      // through 4 KB of swizzled shared memory: written row-per-thread, read back with 8 lanes per row.  In that
      // layout the old x arrives by coalesced 128-bit loads -- issued one chunk ahead, the first chunk while the
      // tile's main loop is still running, so their latency is off the critical path and they do not queue behind
      // the operand TMA loads -- and the new x and its bf16 copy leave by coalesced stores.
      // Stream-K: the unit that holds the TAIL k-range of a split tile (always its first segment) reduce-adds its raw
      // partial into x (TMA) and raises a flag; the unit that holds the HEAD (always its last segment) waits for the
      // flag before it reads x, so the fp32 addition order is fixed: (x + tail) + (head + bias).
      static_assert(CW == 128, "row-square partials are per 128-column group");
      const uint32_t xs = stg;
      const int row0w = q * 32;
      const int lr = lane >> 3, lc = (lane & 7) * 4;  // coalesced layout: row 4 i + lr (i = 0..7), columns lc .. lc + 3 of the chunk
      float* xg = reinterpret_cast<float*>(p.out);
      auto ld_x = [&](float4 (&dst)[8], int row0, int col) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = row0 + 4 * i + lr;
          dst[i] = (r < p.m && col + lc + 4 <= p.n) ? ld_global_nc_f4(xg + static_cast<size_t>(r) * p.ldo + col + lc)
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      for (; work.next(sg); ++it) {
        const int row0 = (sg.tile / num_n) * MT + rank * BM + row0w;
        const int col0 = (sg.tile % num_n) * BN + g * CW;
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        const bool tail = sg.kb0 != 0;
        float4 xr[8];
        if (!tail) {
          if (sg.kb1 != num_kb) {  // head of a split tile: the tail partial of unit + 1 must have landed in x
            if (lane == 0) {
              int32_t* flag = p.sync + ((unit + 1) * NCTA + rank) * (4 * EG) + (warp - 4);
              const long long t0 = clock64();
              while (ld_acquire_gpu(flag) == 0) {
                if (clock64() - t0 > 4000000000ll) {
                  printf("b200vit: stream-K hand-over timed out (block %d warp %d)\n", blockIdx.x, warp);
                  __trap();
                }
              }
              *flag = 0;  // ready for the next launch
            }
            __syncwarp();
          }
          ld_x(xr, row0, col0);  // in flight while the tile's main loop finishes
        }
        float4 bb = bias4(p.bias, col0 + lc, p.n);  // likewise: an L2 round trip per chunk otherwise (ncu: top stall)
        ESTAMP(e_busy);
        mbar_wait(&tfull[as], aph);
        tc_fence_after();
        ESTAMP(e_wait);
#ifdef B200_GEMM_TIMING
        e_lastfull = e_prev;
#endif
        const uint32_t taddr = tmem_base + as * C::ACC_STRIDE + (static_cast<uint32_t>(row0w) << 16) + g * CW;
        if (tail) {
#pragma unroll 1
          for (int c = 0; c < CW; c += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + c, v);
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) st_shared_v4(swz128(xs, lane, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_2d(&tma_out, xs, col0 + c, row0);
              bulk_commit();
            }
          }
          if (lane == 0) bulk_wait_read<0>();  // the staging buffer is free for the next tile's generic-proxy writes
          __syncwarp();
        } else {
          float ss[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) ss[i] = 0.f;
#pragma unroll 1
          for (int c = 0; c < CW; c += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + c, v);
            const int col = col0 + c;
            float4 xn[8];
            float4 bn = bb;
            if (c + 32 < CW) {
              ld_x(xn, row0, col + 32);  // next chunk's old x and bias
              bn = bias4(p.bias, col + 32 + lc, p.n);
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) st_shared_v4(swz128(xs, lane, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + lr;
              const float4 a = ld_shared_f4(swz128(xs, rr, lane & 7));
              const float x0 = xr[i].x + (a.x + bb.x), x1 = xr[i].y + (a.y + bb.y);
              const float x2 = xr[i].z + (a.z + bb.z), x3 = xr[i].w + (a.w + bb.w);
              ss[i] += x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3;
              const int r = row0 + rr;
              if (r < p.m && col + lc + 4 <= p.n) {
                const size_t off = static_cast<size_t>(r) * p.ldo + col + lc;
                *reinterpret_cast<float4*>(xg + off) = make_float4(x0, x1, x2, x3);
                *reinterpret_cast<uint2*>(p.out_bf16 + off) = make_uint2(pack_bf16x2(x0, x1), pack_bf16x2(x2, x3));
              }
            }
            __syncwarp();  // every lane has read the chunk back before the next one overwrites it
            if (c + 32 < CW) {
#pragma unroll
              for (int i = 0; i < 8; ++i) xr[i] = xn[i];
              bb = bn;
            }
          }
          // row sums: the 8 lanes that share a row combine their column partials in a fixed tree
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float s8 = ss[i];
            s8 += __shfl_xor_sync(0xffffffffu, s8, 1);
            s8 += __shfl_xor_sync(0xffffffffu, s8, 2);
            s8 += __shfl_xor_sync(0xffffffffu, s8, 4);
            const int r = row0 + 4 * i + lr;
            if ((lane & 7) == 0 && p.rowsq_out != nullptr && r < p.m && col0 < p.n)
              p.rowsq_out[static_cast<size_t>(col0 / CW) * p.m + r] = s8;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[as]), 0));
          else mbar_arrive(&tempty[as]);
          if (tail) {  // partial landed in x -> hand the tile over to the unit that holds its head
            bulk_wait<0>();
            fence_proxy_async_all();
            st_release_gpu(p.sync + (unit * NCTA + rank) * (4 * EG) + (warp - 4), 1);
          }
        }
      }
    } else
    for (; work.next(sg); ++it) {
      const int m0 = (sg.tile / num_n) * MT + rank * BM;
      const int n0 = (sg.tile % num_n) * BN;
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      float rs = 1.0f;  // the row's RMSNorm scale, fetched while the tile's main loop is still running
      if constexpr (EPI == B200VIT_EPI_QKV_ROPE || EPI == B200VIT_EPI_SWIGLU) {
        const int row = m0 + q * 32 + lane;
        rs = row_rstd(p, row < p.m ? row : 0);
      }
      ESTAMP(e_busy);
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      ESTAMP(e_wait);
      const uint32_t taddr = tmem_base + as * C::ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16) + g * CW;
      epilogue_tile<EPI, CW>(taddr, m0 + q * 32, lane, n0 + g * CW, p, &tma_out, stg, sg.kb0 == 0, rs);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[as]), 0));
        else mbar_arrive(&tempty[as]);
      }
    }
    ESTAMP(e_busy);
    if (epi_stage_bytes(EPI) > 0 && lane == 0) bulk_wait<0>();  // all TMA stores / reductions have landed
#ifdef B200_GEMM_TIMING
    if (blockIdx.x == 10 && warp == 4 && lane == 0) {
      const long long tail = clock64() - e_prev;
      printf("gemm epilogue warp: total %lld | wait tfull %lld | busy %lld | final bulk wait %lld | last tile epilogue %lld | fused attention: staging %lld barriers %lld attention %lld\n", clock64() - e_start, e_wait, e_busy, tail, e_lastfull ? clock64() - e_lastfull : 0, e_a, e_b, e_c);
    }
#endif
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, C::TMEM_COLS); else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

bool stream_k_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200VIT_STREAM_K");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

bool winattn_gate_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200VIT_WINATTN_GATE");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

bool weight_prefetch_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200VIT_BPREFETCH");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

bool use_pair() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200VIT_GEMM_PAIR");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

// How many units (CTAs, or CTA pairs) of `kern` can be resident on the current device at the same time.  Queried once per
// call-site preparation (memoised with the tensor maps); a failed query counts as "enough" (plain devices always fit one
// CTA per SM, which is what the grids are sized for).
template <typename K>
int coresident_units(K kern, int threads, int smem_bytes, bool pair, int sms) {
  if (pair) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(static_cast<unsigned>(sms / 2 * 2));
    cfg.blockDim = dim3(static_cast<unsigned>(threads));
    cfg.dynamicSmemBytes = static_cast<size_t>(smem_bytes);
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, reinterpret_cast<const void*>(kern), &cfg) != cudaSuccess) {
      cudaGetLastError();
      return sms;
    }
    return clusters;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, static_cast<size_t>(smem_bytes)) != cudaSuccess) {
    cudaGetLastError();
    return sms;
  }
  return per_sm * sms;
}

template <int BN, int EG, int EPI, bool PAIR>
int launch_cfg(const b200vit_gemm_args& a, cudaStream_t stream, GemmPrepared* cache) {
  using C = TileCfg<BN, EG, EPI, PAIR>;
  constexpr int MT = PAIR ? 2 * BM : BM;
  GemmPrepared local;
  GemmPrepared& g = cache ? *cache : local;
  auto kern = gemm_tcgen05_kernel<BN, EG, EPI, PAIR>;
  static DeviceOnce attr_set;  // per instantiation and device
  if (attr_set.need()) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set.mark();
  }
  if (!(g.valid && std::memcmp(&g.key, &a, sizeof(a)) == 0)) {
    g.valid = false;
    int rc = make_tmap_bf16(&g.ta, a.d_a, a.m, a.k, BM);
    if (rc) return rc;
    rc = make_tmap_bf16(&g.tb, a.d_b, a.n, a.k, C::BN_LOAD);
    if (rc) return rc;
    g.to = g.ta;
    g.taux = g.ta;
    if (is_resid_norm(EPI)) {
      rc = make_tmap_2d(&g.to, a.d_out, a.m, a.n, a.ldo, 4, 32, 32, 128);  // reduce-add of stream-K tail partials
    } else if (EPI == B200VIT_EPI_QKV_ROPE_WINATTN) {
      rc = make_tmap_2d(&g.to, a.d_out, a.m, a.n / 3, a.ldo, 2, 32, 80, 0);  // attention output [M, D], 32-row boxes
    } else if (EPI == B200VIT_EPI_QKV_ROPE) rc = make_tmap_2d(&g.to, a.d_out, a.m, a.n, a.ldo, 2, 32, 80, 0);
    else if (EPI == B200VIT_EPI_BIAS_RESIDUAL) rc = make_tmap_2d(&g.to, a.d_out, a.m, a.n, a.ldo, 4, 32, 32, 128);
    else if (EPI == B200VIT_EPI_SWIGLU) rc = make_tmap_2d(&g.to, a.d_out, a.m, a.n / 2, a.ldo, 2, 32, 64, 128);
    else if (EPI == B200VIT_EPI_BIAS_GELU) rc = make_tmap_2d(&g.to, a.d_out, a.m, a.n, a.ldo, 2, 32, 64, 128);
    if (rc) return rc;
    const int num_tiles = ((a.m + MT - 1) / MT) * ((a.n + BN - 1) / BN);
    const int sms = device_sm_count();
    int max_units = PAIR ? sms / 2 : sms;
    if (const char* e = getenv("B200VIT_GEMM_MAX_UNITS")) {  // experiment: run on fewer SMs
      const int lim = atoi(e);
      if (lim > 0 && lim < max_units) max_units = lim;
    }
    int units = num_tiles < max_units ? num_tiles : max_units;
    g.stream_k = 0;
    if (is_resid_norm(EPI) && a.d_sync != nullptr && stream_k_enabled()) {
      // balance the k-blocks over all pairs when the tile count is not a multiple of the pair count.  Every unit's
      // range must be longer than one tile (+ the boundary snapping slack), so a tile is cut at most once and the
      // cut is always between the LAST segment of unit u (head) and the FIRST segment of unit u + 1 (tail).
      const int num_kb = (a.k + BK - 1) / BK;
      const long long total = static_cast<long long>(num_tiles) * num_kb;
      // The head of a cut tile waits for the flag of the unit that holds its tail, so every CTA of the grid must be
      // resident at the same time: ask the occupancy calculator (it knows about MPS / green-context SM limits) and fall
      // back to whole tiles when the grid would not fit.  The last 64 ints of the scratch are the full-attention layers'
      // work counters (api.cu).
      if (num_tiles % max_units != 0 && total / max_units >= num_kb + 8 &&
          (max_units + 1) * (PAIR ? 2 : 1) * C::EPI_WARPS <= B200VIT_GEMM_SYNC_INTS - 64 &&
          coresident_units(kern, 128 + 128 * EG, C::SMEM_BYTES, PAIR, sms) >= max_units) {
        units = max_units;
        g.stream_k = 1;
      }
    }
    g.grid = PAIR ? 2 * units : units;
    g.key = a;
    g.valid = true;
  }
  GemmParams p{a.d_out, a.d_bias, a.d_row_map, reinterpret_cast<const float2*>(a.d_rope),
               reinterpret_cast<const int2*>(a.d_rope_pos), a.m, a.n, a.k, a.ldo,
               g.stream_k | (weight_prefetch_enabled() ? 0 : 2) | (winattn_gate_enabled() ? 0 : 4),
               reinterpret_cast<__nv_bfloat16*>(a.d_out_bf16), a.d_rowsq_out, a.d_rowsq_in, a.rowsq_parts, a.norm_eps, a.d_sync};
  B200_CUDA_OK(launch_kernel(kern, dim3(g.grid), dim3(128 + 128 * EG), C::SMEM_BYTES, stream, PAIR ? 2 : 1, g.ta, g.tb, g.to, g.taux, p));
  return 0;
}

template <int BN, int EG, int EPI>
int launch_one(const b200vit_gemm_args& a, cudaStream_t stream, GemmPrepared* cache) {
  if (use_pair()) return launch_cfg<BN, EG, EPI, true>(a, stream, cache);
  return launch_cfg<BN, EG, EPI, false>(a, stream, cache);
}

}  // namespace

int launch_gemm(const b200vit_gemm_args& a, cudaStream_t stream, GemmPrepared* cache) {
  if (a.m <= 0 || a.n <= 0 || a.k <= 0) return fail(B200VIT_EINVAL, "gemm: empty problem");
  if (a.k % 8 != 0) return fail(B200VIT_EINVAL, "gemm: K must be a multiple of 8 (16-byte rows for TMA)");
  if ((reinterpret_cast<uintptr_t>(a.d_a) | reinterpret_cast<uintptr_t>(a.d_b) | reinterpret_cast<uintptr_t>(a.d_out)) & 15)
    return fail(B200VIT_EALIGN, "gemm: A, B and out must be 16-byte aligned");
  const bool needs_bias = a.epilogue != B200VIT_EPI_STORE_F32;
  if (a.d_rowsq_in != nullptr && (a.rowsq_parts <= 0 || a.rowsq_parts > MAX_ROWSQ_PARTS))
    return fail(B200VIT_EINVAL, "gemm: d_rowsq_in needs 1 <= rowsq_parts <= 16 (K <= 2048)");
  if (a.d_rowsq_in != nullptr && a.epilogue != B200VIT_EPI_QKV_ROPE && a.epilogue != B200VIT_EPI_SWIGLU &&
      a.epilogue != B200VIT_EPI_QKV_ROPE_WINATTN)
    return fail(B200VIT_EINVAL, "gemm: d_rowsq_in (fused RMSNorm) is implemented by the QKV_ROPE and SWIGLU epilogues");
  if ((a.d_out_bf16 != nullptr || a.d_rowsq_out != nullptr) && a.epilogue != B200VIT_EPI_STORE_F32 &&
      a.epilogue != B200VIT_EPI_BIAS_RESIDUAL_NORM)
    return fail(B200VIT_EINVAL, "gemm: d_out_bf16 / d_rowsq_out are produced by the STORE_F32 and BIAS_RESIDUAL_NORM epilogues");
  if (needs_bias && a.d_bias == nullptr) return fail(B200VIT_EINVAL, "gemm: epilogue needs a bias vector");
  switch (a.epilogue) {
    case B200VIT_EPI_STORE_F32:
      if (a.n % 8 || a.ldo % 4) return fail(B200VIT_EINVAL, "gemm: N % 8 and ldo % 4 required");
      return launch_one<256, 2, B200VIT_EPI_STORE_F32>(a, stream, cache);
    case B200VIT_EPI_BIAS_F32:
      if (a.n % 8 || a.ldo % 4) return fail(B200VIT_EINVAL, "gemm: N % 8 and ldo % 4 required");
      return launch_one<256, 2, B200VIT_EPI_BIAS_F32>(a, stream, cache);
    case B200VIT_EPI_QKV_ROPE:
      if (a.n % 240 || a.ldo % 8 || !a.d_rope || !a.d_rope_pos ||
          ((reinterpret_cast<uintptr_t>(a.d_rope) | reinterpret_cast<uintptr_t>(a.d_rope_pos)) & 15))
        return fail(B200VIT_EINVAL, "gemm: QKV epilogue needs N % 240 == 0, head_dim 80, a 16-byte aligned rope table and row positions");
      return launch_one<240, 3, B200VIT_EPI_QKV_ROPE>(a, stream, cache);
    case B200VIT_EPI_QKV_ROPE_WINATTN:
      if (a.n % 240 || a.ldo % 8 || a.m % 64 || !a.d_rope || !a.d_rope_pos ||
          ((reinterpret_cast<uintptr_t>(a.d_rope) | reinterpret_cast<uintptr_t>(a.d_rope_pos)) & 15))
        return fail(B200VIT_EINVAL, "gemm: fused window attention needs N % 240 == 0, head_dim 80, M % 64 == 0, a rope table and row positions");
      return launch_one<240, 3, B200VIT_EPI_QKV_ROPE_WINATTN>(a, stream, cache);
    case B200VIT_EPI_BIAS_RESIDUAL:
      if (a.n % 8 || a.ldo % 4) return fail(B200VIT_EINVAL, "gemm: N % 8 and ldo % 4 required");
      return launch_one<256, 2, B200VIT_EPI_BIAS_RESIDUAL>(a, stream, cache);
    case B200VIT_EPI_BIAS_RESIDUAL_NORM:
      if (a.n % 8 || a.ldo % 8 || !a.d_out_bf16 || (reinterpret_cast<uintptr_t>(a.d_out_bf16) & 15))
        return fail(B200VIT_EINVAL, "gemm: BIAS_RESIDUAL_NORM needs N % 8 == 0, ldo % 8 == 0 and a 16-byte aligned d_out_bf16");
      return launch_one<256, 2, B200VIT_EPI_BIAS_RESIDUAL_NORM>(a, stream, cache);
    case B200VIT_EPI_SWIGLU:
      if (a.n % 16 || a.ldo % 8) return fail(B200VIT_EINVAL, "gemm: SwiGLU needs N % 16 == 0 and ldo % 8 == 0");
      return launch_one<256, 2, B200VIT_EPI_SWIGLU>(a, stream, cache);
    case B200VIT_EPI_BIAS_GELU:
      if (a.n % 8 || a.ldo % 8) return fail(B200VIT_EINVAL, "gemm: N % 8 and ldo % 8 required");
      return launch_one<256, 2, B200VIT_EPI_BIAS_GELU>(a, stream, cache);
    case B200VIT_EPI_BIAS_BF16:
      if (a.n % 8 || a.ldo % 8) return fail(B200VIT_EINVAL, "gemm: N % 8 and ldo % 8 required");
      return launch_one<256, 2, B200VIT_EPI_BIAS_BF16>(a, stream, cache);
    default:
      return fail(B200VIT_EINVAL, "gemm: unknown epilogue");
  }
}

}  // namespace b200
