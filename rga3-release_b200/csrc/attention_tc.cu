// tcgen05 / TMEM varlen attention for head_dim 80 (Qwen2.5-VL vision tower, HF
// modeling_qwen2_5_vl.py:244-283; softmax statistics in fp32 as :199).
//
// One CTA = QTILES x 128 consecutive query rows (window order) of one head.  Every row carries the
// bounds [lo, hi) of its own cu_seqlens segment, so the same kernel serves
//   * window layers (QTILES = 1): a 128-row tile holds two 64-patch windows (or several ragged edge
//     windows); the K/V range is the union of their segments and the per-row bounds make the mask
//     block-diagonal; 32-column chunks outside a row's segment are skipped;
//   * full layers (QTILES = 2): 256 rows of one temporal slice (1024 / 2304 patches) share each
//     128-row K/V block; the two query tiles ping-pong on the tensor core while the other tile's
//     softmax runs (softmax is the bound: 128 exp2 per row per block on the 16/clk MUFU).
// Per K/V block and query tile:
//     S = Q K^T   (tcgen05.mma, TMA-loaded operands; head_dim 80 = a 64-wide SWIZZLE_128B sub-tile
//                  + a 16-wide SWIZZLE_32B sub-tile, 5 MMAs of K = 16)
//     P = exp2(S*scale - m)   thread-per-row from TMEM, bf16 into 128-B-swizzled smem
//     O_blk = P V (tcgen05.mma; V in its natural [kv][d] layout as an MN-major B operand)
//     o = o*alpha + O_blk      accumulated in registers (no TMEM rescale pass)
// Warps: 0 = TMA loader, 1 = MMA issuer, 2.. = softmax/accumulate (4 per query tile; TMEM lane
// quadrant = warp % 4).
#include <cuda_bf16.h>

#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "internal.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace b200 {
namespace {

constexpr int HD = 80;
constexpr int QT = 128;   // query rows per tile
constexpr int KVB = 128;  // kv rows per block
constexpr int T64_BYTES = 128 * 128;  // [128 rows][64 cols] bf16
constexpr int T16_BYTES = 128 * 32;   // [128 rows][16 cols] bf16
constexpr int TILE_BYTES = T64_BYTES + T16_BYTES;
constexpr int P_BYTES = 2 * T64_BYTES;  // [128][128] bf16 as two 64-wide K-major sub-tiles
// V is the MN-major B operand of P.V.  Every tcgen05.mma of this shape costs >= 64 cycles whatever its N (the 4 KB A
// operand is re-read from shared memory), so the 80 output columns must come from ONE N=80 instruction per k-step, not a
// 64-wide plus a 16-wide one: the 16-column tail is therefore loaded as a second full 64-column SWIZZLE_128B atom
// (columns 64..127 of the head; the 48 surplus columns belong to the next head or are zero-filled past the matrix edge
// and are never read by the MMA), LBO = T64_BYTES apart.
constexpr int V_TILE_BYTES = 2 * T64_BYTES;

// SHARED_KV = true  (full layers): the QTILES query tiles are consecutive 128-row tiles of ONE head and share
//                     each K/V block (NKV-deep ring); the CTA may walk several heads (NQ Q buffers).
// SHARED_KV = false (window layers): the QTILES "tiles" are the SAME 128 rows of DIFFERENT heads -- two fully
//                     independent (Q, K, V) streams, each with its own softmax warpgroup, interleaved on the
//                     one MMA issuer so one stream's softmax hides the other's tensor-core and load latency.
template <int QTILES, int NKV, int NQ, bool SHARED_KV>
struct AttnCfg {
  static constexpr int NQBUF = SHARED_KV ? NQ * QTILES : QTILES;
  static constexpr int NKBUF = SHARED_KV ? NKV : QTILES;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + NQBUF * TILE_BYTES;
  static constexpr int OFF_V = OFF_K + NKBUF * TILE_BYTES;
  static constexpr int OFF_P = OFF_V + NKBUF * V_TILE_BYTES;
  static constexpr int OFF_BAR = OFF_P + QTILES * P_BYTES;
  static constexpr int BYTES = OFF_BAR + 256 + 1024;
  static constexpr int TMEM_COLS = QTILES == 2 ? 512 : 256;
  static constexpr int THREADS = 64 + 128 * QTILES;
  static constexpr int NQBAR = SHARED_KV ? NQ : QTILES;
  __host__ __device__ static constexpr int S_COL(int t) { return t * 128; }
  __host__ __device__ static constexpr int O_COL(int t) { return QTILES * 128 + t * 128; }
};

// ---- shared-memory operand descriptors (see ptx.cuh::umma_desc_k128 for the field layout)
__device__ __forceinline__ uint64_t desc_common(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= layout << 61;
  return d;
}
constexpr uint64_t SW128 = 2, SW32 = 6;
// K-major, 32-byte rows (16 bf16), SWIZZLE_32B: 8-row groups 256 B apart
__device__ __forceinline__ uint64_t desc_k_sw32(uint32_t saddr) { return desc_common(saddr, 16, 256, SW32); }
// MN-major B operand, rows = K index: 128-byte rows (64 bf16 of N), 8-row groups 1024 B apart
// and 64-column atoms T64_BYTES apart (canonical ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)) in elements)
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr) { return desc_common(saddr, T64_BYTES, 1024, SW128); }

__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int QTILES, int NKV, int NQ, bool SHARED_KV>
__global__ void __launch_bounds__(AttnCfg<QTILES, NKV, NQ, SHARED_KV>::THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm16,
               const __grid_constant__ CUtensorMap to64, const __grid_constant__ CUtensorMap to16,
               const AttnTile* __restrict__ tiles, const int2* __restrict__ bounds, int m_rows, int heads,
               int heads_per_cta, float scale_log2) {
  using L = AttnCfg<QTILES, NKV, NQ, SHARED_KV>;
  griddep_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* q_full = bars;  // [NQBAR]
  uint64_t* q_empty = q_full + L::NQBAR;
  uint64_t* k_full = q_empty + L::NQBAR;  // [NKBUF]
  uint64_t* k_empty = k_full + L::NKBUF;
  uint64_t* v_full = k_empty + L::NKBUF;
  uint64_t* v_empty = v_full + L::NKBUF;
  uint64_t* s_full = v_empty + L::NKBUF;  // [QTILES]
  uint64_t* p_full = s_full + QTILES;
  uint64_t* o_full = p_full + QTILES;
  uint64_t* o_empty = o_full + QTILES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + QTILES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const AttnTile tile = tiles[blockIdx.x];
  const int head0 = blockIdx.y * heads_per_cta;
  const int D = heads * HD;
  const int nblk = tile.n_kv_blocks;
  // head iterations per stream and flattened (head, kv block) iterations
  const int n_hl = SHARED_KV ? heads_per_cta : heads_per_cta / QTILES;
  const int n_iter = n_hl * nblk;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm64);
    tma_prefetch_desc(&tm16);
    tma_prefetch_desc(&to64);
    tma_prefetch_desc(&to16);
    for (int b = 0; b < L::NQBAR; ++b) {
      mbar_init(&q_full[b], 1);
      mbar_init(&q_empty[b], 1);
    }
    for (int b = 0; b < L::NKBUF; ++b) {
      mbar_init(&k_full[b], 1);
      mbar_init(&k_empty[b], 1);
      mbar_init(&v_full[b], 1);
      mbar_init(&v_empty[b], 1);
    }
    for (int t = 0; t < QTILES; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&o_full[t], 1);
      mbar_init(&o_empty[t], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, L::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA loader =====================
      auto load_tile = [&](uint8_t* dst, uint64_t* bar, int col, int row) {
        tma_load_2d(dst, &tm64, bar, col, row);
        tma_load_2d(dst + T64_BYTES, &tm16, bar, col + 64, row);
      };
      auto load_v = [&](uint8_t* dst, uint64_t* bar, int col, int row) {
        tma_load_2d(dst, &tm64, bar, col, row);
        tma_load_2d(dst + T64_BYTES, &tm64, bar, col + 64, row);
      };
      if constexpr (SHARED_KV) {
        int i = 0;
        for (int hl = 0; hl < n_hl; ++hl) {
          const int head = head0 + hl;
          const int qb = hl % NQ;
          mbar_wait(&q_empty[qb], ((hl / NQ) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[qb], QTILES * TILE_BYTES);
#pragma unroll
          for (int t = 0; t < QTILES; ++t)
            load_tile(smem + L::OFF_Q + (qb * QTILES + t) * TILE_BYTES, &q_full[qb], head * HD, tile.q_row0 + t * QT);
          for (int j = 0; j < nblk; ++j, ++i) {
            const int b = i % NKV;
            const uint32_t par = ((i / NKV) & 1) ^ 1;
            const int row = tile.kv_row0 + j * KVB;
            mbar_wait(&k_empty[b], par);
            mbar_arrive_expect_tx(&k_full[b], TILE_BYTES);
            load_tile(smem + L::OFF_K + b * TILE_BYTES, &k_full[b], D + head * HD, row);
            mbar_wait(&v_empty[b], par);
            mbar_arrive_expect_tx(&v_full[b], V_TILE_BYTES);
            load_v(smem + L::OFF_V + b * V_TILE_BYTES, &v_full[b], 2 * D + head * HD, row);
          }
        }
      } else {
        // Independent streams, single-buffered Q/K/V each: issue whichever buffer has been released instead of walking
        // the streams in a fixed order (a stream waiting for its P.V to free V must not hold back the other stream's
        // next Q/K, which were free since that stream's Q.K^T).  Per stream: it = hl * nblk + j.
        const int n_it = n_hl * nblk;
        int next_k[QTILES], next_v[QTILES];
#pragma unroll
        for (int t = 0; t < QTILES; ++t) next_k[t] = next_v[t] = 0;
        int remaining = 2 * QTILES * n_it;
        const long long t0 = clock64();
        while (remaining > 0) {
          bool progressed = false;
#pragma unroll
          for (int t = 0; t < QTILES; ++t) {
            int it = next_k[t];
            if (it < n_it) {
              const int hl = it / nblk, j = it - hl * nblk;
              const int head = head0 + hl * QTILES + t;
              const bool need_q = j == 0;
              if (mbar_test(&k_empty[t], (it & 1) ^ 1) && (!need_q || mbar_test(&q_empty[t], (hl & 1) ^ 1))) {
                if (need_q) {
                  mbar_arrive_expect_tx(&q_full[t], TILE_BYTES);
                  load_tile(smem + L::OFF_Q + t * TILE_BYTES, &q_full[t], head * HD, tile.q_row0);
                }
                mbar_arrive_expect_tx(&k_full[t], TILE_BYTES);
                load_tile(smem + L::OFF_K + t * TILE_BYTES, &k_full[t], D + head * HD, tile.kv_row0 + j * KVB);
                next_k[t] = it + 1;
                --remaining;
                progressed = true;
              }
            }
            it = next_v[t];
            if (it < n_it && mbar_test(&v_empty[t], (it & 1) ^ 1)) {
              const int hl = it / nblk, j = it - hl * nblk;
              const int head = head0 + hl * QTILES + t;
              mbar_arrive_expect_tx(&v_full[t], V_TILE_BYTES);
              load_v(smem + L::OFF_V + t * V_TILE_BYTES, &v_full[t], 2 * D + head * HD, tile.kv_row0 + j * KVB);
              next_v[t] = it + 1;
              --remaining;
              progressed = true;
            }
          }
          if (!progressed) {
            __nanosleep(40);
            if (clock64() - t0 > 8000000000ll) {
              printf("b200vit: attention loader timed out (block %d,%d)\n", blockIdx.x, blockIdx.y);
              __trap();
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_qk = idesc_bf16(QT, KVB, false);
      constexpr uint32_t idesc_pv = idesc_bf16(QT, HD, true);
      auto kv_stage = [&](int t, int i) { return SHARED_KV ? (i % NKV) : t; };
      auto q_slot = [&](int t, int i) { return SHARED_KV ? (((i / nblk) % NQ) * QTILES + t) : t; };
      auto issue_qk = [&](int t, int i) {  // S(t) is free: the caller has waited p_full(t, i-1)
        const uint32_t q64 = smem_u32(smem + L::OFF_Q + q_slot(t, i) * TILE_BYTES), q16 = q64 + T64_BYTES;
        const uint32_t k64 = smem_u32(smem + L::OFF_K + kv_stage(t, i) * TILE_BYTES), k16 = k64 + T64_BYTES;
        const uint32_t ts = tmem_base + L::S_COL(t);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(ts, umma_desc_k128(q64 + k * 32), umma_desc_k128(k64 + k * 32), idesc_qk, k != 0 ? 1u : 0u);
        umma_bf16_ss(ts, desc_k_sw32(q16), desc_k_sw32(k16), idesc_qk, 1u);
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int i) {
        mbar_wait(&p_full[t], i & 1);         // softmax(t, i) has written P(t) and released S(t)
        mbar_wait(&o_empty[t], (i & 1) ^ 1);  // accumulate(t, i-1) has drained O(t)
        tc_fence_after();
        const uint32_t p0 = smem_u32(smem + L::OFF_P + t * P_BYTES);
        const uint32_t v64 = smem_u32(smem + L::OFF_V + kv_stage(t, i) * V_TILE_BYTES);
        const uint32_t to = tmem_base + L::O_COL(t);
#pragma unroll
        for (int k = 0; k < KVB / 16; ++k) {
          const uint64_t pa = umma_desc_k128(p0 + (k >> 2) * T64_BYTES + (k & 3) * 32);
          umma_bf16_ss(to, pa, desc_mn_sw128(v64 + k * 2048), idesc_pv, k != 0 ? 1u : 0u);
        }
        umma_commit(&o_full[t]);
      };
      if constexpr (SHARED_KV) {
        mbar_wait(&q_full[0], 0);
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < QTILES; ++t) issue_qk(t, 0);
        umma_commit(&k_empty[0]);
        if (nblk == 1) umma_commit(&q_empty[0]);
        for (int i = 0; i < n_iter; ++i) {
          const int b = i % NKV;
          mbar_wait(&v_full[b], (i / NKV) & 1);
          const bool more = i + 1 < n_iter;
          if (more) {
            mbar_wait(&k_full[(i + 1) % NKV], ((i + 1) / NKV) & 1);
            if ((i + 1) % nblk == 0) {  // first block of the next head: its Q must have landed
              const int hl = (i + 1) / nblk;
              mbar_wait(&q_full[hl % NQ], (hl / NQ) & 1);
            }
          }
#pragma unroll
          for (int t = 0; t < QTILES; ++t) {
            issue_pv(t, i);
            if (more) issue_qk(t, i + 1);
          }
          umma_commit(&v_empty[b]);
          if (more) {
            umma_commit(&k_empty[(i + 1) % NKV]);
            if ((i + 1) % nblk == nblk - 1) umma_commit(&q_empty[((i + 1) / nblk) % NQ]);  // last Q K^T of that head
          }
        }
      } else {
        // Independent streams: issue whichever stream's next operation is ready (a blocked stream must not
        // hold up the other one's P V).  Per stream the order is QK(0) PV(0) QK(1) PV(1) ...
        int next_op[QTILES];  // 2*i = QK(i), 2*i+1 = PV(i)
#pragma unroll
        for (int t = 0; t < QTILES; ++t) next_op[t] = 0;
        int remaining = QTILES * 2 * n_iter;
        const long long t0 = clock64();
        while (remaining > 0) {
          bool progressed = false;
#pragma unroll
          for (int t = 0; t < QTILES; ++t) {
            const int op = next_op[t];
            if (op >= 2 * n_iter) continue;
            const int i = op >> 1;
            if ((op & 1) == 0) {  // QK(i): operands landed (S(t) is free: PV(i-1) was issued after p_full(i-1))
              if (!mbar_test(&k_full[t], i & 1)) continue;
              if (i % nblk == 0 && !mbar_test(&q_full[t], (i / nblk) & 1)) continue;
              tc_fence_after();
              issue_qk(t, i);
              umma_commit(&k_empty[t]);
              if (i % nblk == nblk - 1) umma_commit(&q_empty[t]);
            } else {              // PV(i): V landed, softmax done, O(t) drained
              // p_full is the one that is normally still pending: test it first so an idle poll costs one probe
              if (!mbar_test(&p_full[t], i & 1) || !mbar_test(&v_full[t], i & 1) || !mbar_test(&o_empty[t], (i & 1) ^ 1))
                continue;
              issue_pv(t, i);  // its own waits succeed immediately
              umma_commit(&v_empty[t]);
            }
            next_op[t] = op + 1;
            --remaining;
            progressed = true;
          }
          if (!progressed) __nanosleep(40);  // do not steal issue slots from the softmax warps on this sub-partition
          if (!progressed && clock64() - t0 > 8000000000ll) {
            printf("b200vit: attention scheduler timed out (block %d,%d)\n", blockIdx.x, blockIdx.y);
            __trap();
          }
        }
      }
    }
  } else {
    // ===================== softmax + output accumulation: one thread per query row =====================
    const int t = (warp - 2) >> 2;   // query tile of this warp
    const int quad = warp & 3;       // TMEM lane quadrant
    const int r = quad * 32 + lane;  // row inside the tile == TMEM lane
    const int row = tile.q_row0 + (SHARED_KV ? t * QT : 0) + r;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t ts = lane_base + L::S_COL(t);
    const uint32_t to = lane_base + L::O_COL(t);
    const uint32_t pbuf = smem_u32(smem + L::OFF_P + t * P_BYTES);
    int2 bd = make_int2(0, 0);
    if (row < m_rows) bd = bounds[row];
    float o[HD];
    float m_run, l_run, alpha_prev;

    auto accumulate_block = [&](int i, float alpha) {
      mbar_wait(&o_full[t], i & 1);
      tc_fence_after();
      {
        uint32_t v[48];  // 48 + 32 columns: two TMEM round trips instead of five
        tmem_ld32(to, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld16(to + 32, *reinterpret_cast<uint32_t(*)[16]>(&v[32]));
        tmem_ld_wait();
#pragma unroll
        for (int i2 = 0; i2 < 48; ++i2) o[i2] = o[i2] * alpha + __uint_as_float(v[i2]);
        tmem_ld32(to + 48, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld_wait();
#pragma unroll
        for (int i2 = 0; i2 < 32; ++i2) o[48 + i2] = o[48 + i2] * alpha + __uint_as_float(v[i2]);
      }
      tc_fence_before();
      mbar_arrive(&o_empty[t]);
    };

#ifdef B200_ATTN_TIMING
    long long tstamp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
    const bool trec = (blockIdx.x == 1 && blockIdx.y == 0 && warp == 2 && lane == 0);
#define TSTAMP(k) do { long long _n = clock64(); tstamp[k] += _n - tprev; tprev = _n; } while (0)
#else
#define TSTAMP(k)
#endif
    for (int hl = 0; hl < n_hl; ++hl) {
    m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
#pragma unroll
    for (int i2 = 0; i2 < HD; ++i2) o[i2] = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int it = hl * nblk + j;  // flattened iteration (barrier phases)
      TSTAMP(0);
      mbar_wait(&s_full[t], it & 1);
      tc_fence_after();
      TSTAMP(1);
      // valid kv range of this row inside the block, in columns [0, 128)
      const int kv0 = tile.kv_row0 + j * KVB;
      const int lo = max(bd.x - kv0, 0), hi = min(bd.y - kv0, KVB);
      // tcgen05.ld is warp-collective: skip / fast-path decisions are made per WARP (ballots), the
      // per-row bounds only enter through the masked path.
      // ---- pass 1: row maximum over the valid columns (64 columns per TMEM round trip)
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < KVB; c += 64) {
        if (__all_sync(0xffffffffu, c + 64 <= lo || c >= hi)) continue;  // outside every row's segment
        uint32_t v[64];
        tmem_ld32(ts + c, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32(ts + c + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld_wait();
        if (__all_sync(0xffffffffu, c >= lo && c + 64 <= hi)) {
          float m4[4] = {mx, -INFINITY, -INFINITY, -INFINITY};  // independent chains: a single fmax chain is latency-bound
#pragma unroll
          for (int i = 0; i < 64; i += 4) {
            m4[0] = fmaxf(m4[0], __uint_as_float(v[i]));
            m4[1] = fmaxf(m4[1], __uint_as_float(v[i + 1]));
            m4[2] = fmaxf(m4[2], __uint_as_float(v[i + 2]));
            m4[3] = fmaxf(m4[3], __uint_as_float(v[i + 3]));
          }
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) mx = fmaxf(mx, (c + i >= lo && c + i < hi) ? __uint_as_float(v[i]) : -INFINITY);
        }
      }
      TSTAMP(2);
      const float m_new = fmaxf(m_run, mx * scale_log2);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;  // nothing valid so far: p = 0, no NaN
      const float alpha = ex2_approx(m_run - m_use);           // m_run = -inf -> 0
      // ---- pass 2: probabilities -> bf16 -> swizzled smem (A operand of P V), 32 columns per trip
      if (lane == 0) bulk_wait_read<0>();  // the previous head's output store (staged in this P region) has drained
      __syncwarp();
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < KVB; c += 32) {
        const uint32_t sub = pbuf + (c >> 6) * T64_BYTES;  // 64-wide sub-tile
        const int j0 = (c & 63) >> 3;                      // first 16-byte chunk of this 32-column group
        if (__all_sync(0xffffffffu, c + 32 <= lo || c >= hi)) {
#pragma unroll
          for (int q = 0; q < 4; ++q) st_shared_v4(swz128(sub, r, j0 + q), 0u, 0u, 0u, 0u);
          continue;
        }
        uint32_t v[32];
        tmem_ld32(ts + c, v);
        tmem_ld_wait();
        uint32_t pk[16];
        if (__all_sync(0xffffffffu, c >= lo && c + 32 <= hi)) {
          float s4[4] = {0.f, 0.f, 0.f, 0.f};  // independent partial sums (ILP)
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float p0 = ex2_approx(__uint_as_float(v[i]) * scale_log2 - m_use);
            const float p1 = ex2_approx(__uint_as_float(v[i + 1]) * scale_log2 - m_use);
            const float p2 = ex2_approx(__uint_as_float(v[i + 2]) * scale_log2 - m_use);
            const float p3 = ex2_approx(__uint_as_float(v[i + 3]) * scale_log2 - m_use);
            s4[0] += p0, s4[1] += p1, s4[2] += p2, s4[3] += p3;
            pk[i >> 1] = pack_bf16x2(p0, p1);
            pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
          }
          sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = ex2_approx(__uint_as_float(v[i]) * scale_log2 - m_use);
            float p1 = ex2_approx(__uint_as_float(v[i + 1]) * scale_log2 - m_use);
            p0 = (c + i >= lo && c + i < hi) ? p0 : 0.f;
            p1 = (c + i + 1 >= lo && c + i + 1 < hi) ? p1 : 0.f;
            sum += p0 + p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          st_shared_v4(swz128(sub, r, j0 + q), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
      l_run = l_run * alpha + sum;
      m_run = m_new;
      TSTAMP(3);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      TSTAMP(4);
      // the previous block's P V has long finished: fold it in while the tensor core works on this one
      if (j >= 1) accumulate_block(it - 1, alpha_prev);
      alpha_prev = alpha;
    }
    accumulate_block(hl * nblk + nblk - 1, alpha_prev);
    TSTAMP(5);

    {
      // O -> bf16 -> this warp's own rows of the (now idle) P buffer -> two TMA stores (64 + 16 columns).
      // A thread owns a row, so direct global stores would be 32 scattered sectors per instruction.
      const float inv = (l_run > 0.f) ? 1.f / l_run : 0.f;
      const int head = SHARED_KV ? head0 + hl : head0 + hl * QTILES + t;
      const uint32_t stg64 = pbuf + quad * 4096;               // rows quad*32.. of sub-tile 0 (1024-aligned)
      const uint32_t stg16 = pbuf + T64_BYTES + quad * 4096;   // same rows of sub-tile 1
#pragma unroll
      for (int c = 0; c < 64; c += 8)
        st_shared_v4(swz128(stg64, lane, c >> 3), pack_bf16x2(o[c] * inv, o[c + 1] * inv), pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv),
                     pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv), pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv));
#pragma unroll
      for (int c = 64; c < HD; c += 8)
        st_shared_v4(stg16 + lane * 32 + (c - 64) * 2, pack_bf16x2(o[c] * inv, o[c + 1] * inv),
                     pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv), pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv),
                     pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const int row0 = tile.q_row0 + (SHARED_KV ? t * QT : 0) + quad * 32;
        tma_store_2d(&to64, stg64, head * HD, row0);        // rows >= m_rows are clipped by the tensor map
        tma_store_2d(&to16, stg16, head * HD + 64, row0);
        bulk_commit();
      }
    }
    TSTAMP(6);
    }  // heads of this CTA
    if (lane == 0) bulk_wait<0>();  // output stores have landed before the CTA exits
#ifdef B200_ATTN_TIMING
    if (trec) printf("attn timing (cycles, %d head-iters): other %lld | wait_s %lld | pass1 %lld | pass2 %lld | fence+arrive %lld | wait_o+acc %lld | store %lld\n",
                     n_hl, tstamp[0], tstamp[1], tstamp[2], tstamp[3], tstamp[4], tstamp[5], tstamp[6]);
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ full-attention layers (persistent)
// Two 128-row query tiles of one head share each K/V block (as attn_tc_kernel<2, 2, 1, true>), one thread per query row,
// with changes that each came out of a measurement:
//  * O stays in TMEM for the whole K/V walk: P.V(j) accumulates onto P.V(j-1) and the softmax warps never read a
//    per-block O back.  The running maximum is therefore LAZY: a row keeps the maximum m it has used so far while the
//    new block's maximum stays below m + 8 (log2 units; P <= 2^8, harmless in bf16 / fp32), and only when it grows past
//    that does the row's thread rescale its O columns in TMEM (tcgen05.ld -> * 2^(m - m') -> tcgen05.st) and its sum.
//    After the first block that is rare, so the per-block chain of a tile is  S -> row max -> exp2 pass -> P.V  without
//    the O drain (80 FFMA per row) and without the O hand-shake.
//  * a block's 128 scores are read from TMEM ONCE into registers and S(t) is released at once (s_free), so the next
//    Q.K^T runs underneath the softmax and S(j+1) is waiting when the warps come back; P.V(j-1) is only waited for
//    after the whole exp2 pass, just before P is overwritten.
//  * one MMA-issuing warp per tile on blocking (hardware-suspended) mbarrier waits.  (A single issuer polling both tiles
//    with nanosleep between probes reacted late: the sleep quantum is ~1 us.)
//  * PERSISTENT: one CTA per SM walks work items (256-row tile pair, head) -- the first is blockIdx.x, the following
//    ones are drawn from a device counter (or blockIdx.x + k * gridDim.x without one) -- tile index fastest
//    (co-resident CTAs share a head's K/V in L2).  Barrier phases run on global block / item counters, so
//    the K/V ring, S and P flow straight across item boundaries: the next item's Q is loaded as soon as the last Q.K^T
//    of the current one has been issued, its first Q.K^T runs underneath the current item's last softmax, and an item's
//    output store drains while the next item's first block is computed.
constexpr float RESCALE_LOG2 = 8.f;
constexpr int F_THREADS = 96 + 256;  // loader, two MMA issuers, 2 x 4 softmax warps
__global__ void __launch_bounds__(F_THREADS, 1)
attn_full_kernel(const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm16,
                  const __grid_constant__ CUtensorMap to64, const __grid_constant__ CUtensorMap to16,
                  const AttnTile* __restrict__ tiles, int n_tiles, const int2* __restrict__ bounds, int m_rows, int heads,
                  float scale_log2, int mufu_token, int32_t* __restrict__ work_counter) {
  using L = AttnCfg<2, 2, 1, true>;
  constexpr int NKV = 2;
  griddep_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 1;   // both tiles' last Q.K^T of an item have finished reading Q
  uint64_t* k_full = bars + 2;
  uint64_t* k_empty = k_full + NKV;
  uint64_t* v_full = k_empty + NKV;
  uint64_t* v_empty = v_full + NKV;
  uint64_t* s_full = v_empty + NKV;  // [2]
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* s_free = o_full + 2;  // [2] softmax(t, g) holds S(t) in registers: the next Q.K^T may overwrite it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);
  // the work item of ordinal n is published in item_slot[n & 1] before q_full completes phase n (-1: no more work)
  volatile int32_t* item_slot = reinterpret_cast<volatile int32_t*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HD;
  const int n_items = n_tiles * heads;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm64);
    tma_prefetch_desc(&tm16);
    tma_prefetch_desc(&to64);
    tma_prefetch_desc(&to16);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 2);
    for (int b = 0; b < NKV; ++b) {
      mbar_init(&k_full[b], 1);
      mbar_init(&k_empty[b], 2);  // one commit per tile's issuing warp
      mbar_init(&v_full[b], 1);
      mbar_init(&v_empty[b], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&o_full[t], 1);
      mbar_init(&s_free[t], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA loader =====================
      auto load_tile = [&](uint8_t* dst, uint64_t* bar, int col, int row) {
        tma_load_2d(dst, &tm64, bar, col, row);
        tma_load_2d(dst + T64_BYTES, &tm16, bar, col + 64, row);
      };
      auto load_v = [&](uint8_t* dst, uint64_t* bar, int col, int row) {
        tma_load_2d(dst, &tm64, bar, col, row);
        tma_load_2d(dst + T64_BYTES, &tm64, bar, col + 64, row);
      };
      // Work items: the first one is blockIdx.x, the following ones come from a global counter (work_counter, zero at
      // launch) so that an SM that runs slower, or starts late, simply takes fewer items; without a counter the walk is
      // the static stride.  Only this thread draws items; the other warps read them from item_slot.
      int g = 0, n = 0;  // global K/V block and item ordinals of this CTA
      for (int item = blockIdx.x;; ++n) {
        mbar_wait(q_empty, (n & 1) ^ 1);
        item_slot[n & 1] = item;
        if (item < 0) {
          mbar_arrive(q_full);  // publishes the end marker
          break;
        }
        const AttnTile tile = tiles[item % n_tiles];
        const int head = item / n_tiles;
        mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
        for (int t = 0; t < 2; ++t) load_tile(smem + L::OFF_Q + t * TILE_BYTES, q_full, head * HD, tile.q_row0 + t * QT);
        int next = work_counter != nullptr ? static_cast<int>(gridDim.x) + atomicAdd(work_counter, 1) : item + static_cast<int>(gridDim.x);
        for (int j = 0; j < tile.n_kv_blocks; ++j, ++g) {
          const int b = g % NKV;
          const uint32_t par = ((g / NKV) & 1) ^ 1;
          const int row = tile.kv_row0 + j * KVB;
          mbar_wait(&k_empty[b], par);
          mbar_arrive_expect_tx(&k_full[b], TILE_BYTES);
          load_tile(smem + L::OFF_K + b * TILE_BYTES, &k_full[b], D + head * HD, row);
          mbar_wait(&v_empty[b], par);
          mbar_arrive_expect_tx(&v_full[b], V_TILE_BYTES);
          load_v(smem + L::OFF_V + b * V_TILE_BYTES, &v_full[b], 2 * D + head * HD, row);
        }
        item = next < n_items ? next : -1;
      }
    }
  } else if (warp <= 2) {
    if (lane == 0) {
      // ===================== MMA issuers: one warp per query tile =====================
      // Each walks its own tile's fixed order   Q.K^T(0);  for every block g: [S(t) read by softmax(g): Q.K^T(g+1)]
      // [P(t, g) written: P.V(g)]   on blocking (hardware-suspended) mbarrier waits, straight across item boundaries.  The
      // softmax warps release S(t) as soon as a block's scores are in their registers, so Q.K^T(g+1) runs underneath
      // softmax(g) and S(g+1) is waiting when they come back.  (A single issuer polling both tiles with nanosleep between
      // probes reacted late: the sleep quantum is ~1 us.)  The tensor pipe interleaves the two tiles' MMAs in arrival
      // order; a K/V stage / the Q buffer is released when BOTH warps have committed their use of it (count 2).
      constexpr uint32_t idesc_qk = idesc_bf16(QT, KVB, false);
      constexpr uint32_t idesc_pv = idesc_bf16(QT, HD, true);
      const int t = warp - 1;
      auto issue_qk = [&](int g) {
        const uint32_t q64 = smem_u32(smem + L::OFF_Q + t * TILE_BYTES), q16 = q64 + T64_BYTES;
        const uint32_t k64 = smem_u32(smem + L::OFF_K + (g % NKV) * TILE_BYTES), k16 = k64 + T64_BYTES;
        const uint32_t ts = tmem_base + L::S_COL(t);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(ts, umma_desc_k128(q64 + k * 32), umma_desc_k128(k64 + k * 32), idesc_qk, k != 0 ? 1u : 0u);
        umma_bf16_ss(ts, desc_k_sw32(q16), desc_k_sw32(k16), idesc_qk, 1u);
        umma_commit(&s_full[t]);
        umma_commit(&k_empty[g % NKV]);
      };
      auto issue_pv = [&](int g, bool first) {  // O(t) (+)= P(t) V(g): the softmax warps rescale O themselves when a row maximum grows
        const uint32_t p0 = smem_u32(smem + L::OFF_P + t * P_BYTES);
        const uint32_t v64 = smem_u32(smem + L::OFF_V + (g % NKV) * V_TILE_BYTES);
        const uint32_t to = tmem_base + L::O_COL(t);
#pragma unroll
        for (int k = 0; k < KVB / 16; ++k) {
          const uint64_t pa = umma_desc_k128(p0 + (k >> 2) * T64_BYTES + (k & 3) * 32);
          umma_bf16_ss(to, pa, desc_mn_sw128(v64 + k * 2048), idesc_pv, (!first || k != 0) ? 1u : 0u);
        }
        umma_commit(&o_full[t]);
        umma_commit(&v_empty[g % NKV]);
      };
      int g = 0, n = 0;
#ifdef B200_ATTN_MMA_PROBE
      long long probe_issue = 0, probe_done = 0;
      int probe_n = 0;
#endif
      mbar_wait(q_full, 0);
      int nblk = tiles[item_slot[0] % n_tiles].n_kv_blocks;  // the first item always exists (grid <= items)
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0);
      if (nblk == 1) umma_commit(q_empty);
      while (nblk > 0) {
        int next_nblk = 0;
        for (int i = 0; i < nblk; ++i, ++g) {
          // the successor block's Q.K^T first: it only needs S(t) back, not P(t, g)
          const bool in_item = i + 1 < nblk;
          if (!in_item) {  // the successor is the next item's first block, if there is a next item
            mbar_wait(q_full, (n + 1) & 1);
            const int next_item = item_slot[(n + 1) & 1];
            next_nblk = next_item >= 0 ? tiles[next_item % n_tiles].n_kv_blocks : 0;
          }
          if (in_item || next_nblk > 0) {
            mbar_wait(&s_free[t], g & 1);
            mbar_wait(&k_full[(g + 1) % NKV], ((g + 1) / NKV) & 1);
            tc_fence_after();
            issue_qk(g + 1);
            if (in_item ? (i + 2 == nblk) : (next_nblk == 1)) umma_commit(q_empty);  // that was the last Q.K^T of its item
          }
          mbar_wait(&v_full[g % NKV], (g / NKV) & 1);
          mbar_wait(&p_full[t], g & 1);
          tc_fence_after();
#ifdef B200_ATTN_MMA_PROBE
          const long long c0 = clock64();
          issue_pv(g, i == 0);
          const long long c1 = clock64();
          mbar_wait(&o_full[t], g & 1);
          const long long c2 = clock64();
          probe_issue += c1 - c0, probe_done += c2 - c0, ++probe_n;
#else
          issue_pv(g, i == 0);
#endif
        }
        nblk = next_nblk;
        ++n;
      }
#ifdef B200_ATTN_MMA_PROBE
      if (blockIdx.x == 1) printf("mma warp %d: P.V issue %lld cycles, issue->complete %lld cycles (avg of %d)\n", t, probe_issue / probe_n, probe_done / probe_n, probe_n);
#endif
    }
  } else {
    // ===================== softmax: one thread per query row =====================
    constexpr int SC = KVB;  // score columns per thread
    constexpr int OC = HD;   // output columns per thread
    const int sw = warp - 3;
    const int t = sw >> 2;          // query tile
    const int quad = warp & 3;      // TMEM lane quadrant
    const int r = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t ts = lane_base + L::S_COL(t);
    const uint32_t to = lane_base + L::O_COL(t);
    const uint32_t pbuf = smem_u32(smem + L::OFF_P + t * P_BYTES);
    const uint32_t stg64 = pbuf + quad * 4096;              // output staging [32 rows x 64 cols], SWIZZLE_128B
    const uint32_t stg16 = pbuf + T64_BYTES + quad * 4096;  // output staging [32 rows x 16 cols], dense
    const int tok_mine = 3 + t * 4 + quad, tok_other = 3 + (t ^ 1) * 4 + quad;  // named barriers: MUFU token (below)
    int g = 0;   // global block ordinal (barrier phases)
#ifdef B200_ATTN_TIMING
    long long tstamp[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
    const bool trec = (blockIdx.x == 1 && (warp == 3 || warp == 7) && lane == 0);
    int n_rescale = 0, n_blocks = 0;
#define TSTAMP2(k) do { long long _n = clock64(); tstamp[k] += _n - tprev; tprev = _n; } while (0)
#else
#define TSTAMP2(k)
#endif

    for (int n = 0;; ++n) {
      mbar_wait(q_full, n & 1);
      const int item = item_slot[n & 1];
      if (item < 0) break;
      const AttnTile tile = tiles[item % n_tiles];
      const int head = item / n_tiles;
      const int nblk = tile.n_kv_blocks;
      const int row = tile.q_row0 + t * QT + r;
      int2 bd = make_int2(0, 0);
      if (row < m_rows) bd = bounds[row];
      float m_used = -INFINITY, l_part = 0.f;  // the maximum the row's probabilities are currently expressed against

      for (int j = 0; j < nblk; ++j, ++g) {
        TSTAMP2(0);
        mbar_wait(&s_full[t], g & 1);
        tc_fence_after();
        TSTAMP2(1);
        const int kv0 = tile.kv_row0 + j * KVB;                        // first kv row of the block
        const int lo = max(bd.x - kv0, 0), hi = min(bd.y - kv0, SC);   // this row's valid columns
        // the S(t) row into registers, then S(t) is free for the next Q.K^T
        uint32_t sv[SC];
#pragma unroll
        for (int c = 0; c < SC; c += 32) tmem_ld32(ts + c, *reinterpret_cast<uint32_t(*)[32]>(&sv[c]));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[t]);
        const bool all_valid = __all_sync(0xffffffffu, lo == 0 && hi == SC);
        float mx;
        if (all_valid) {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // independent chains (fmax3 pairs them up)
#pragma unroll
          for (int i = 0; i < SC; i += 4) {
            m4[0] = fmaxf(m4[0], __uint_as_float(sv[i]));
            m4[1] = fmaxf(m4[1], __uint_as_float(sv[i + 1]));
            m4[2] = fmaxf(m4[2], __uint_as_float(sv[i + 2]));
            m4[3] = fmaxf(m4[3], __uint_as_float(sv[i + 3]));
          }
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        } else {
          mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < SC; ++i) mx = fmaxf(mx, (i >= lo && i < hi) ? __uint_as_float(sv[i]) : -INFINITY);
        }
        TSTAMP2(2);
        // lazy maximum
        const float m_blk = mx * scale_log2;
        const bool grow = m_blk > m_used + RESCALE_LOG2;  // also the first block with a valid column (m_used = -inf)
        float alpha = 1.f;
        if (grow) {
          alpha = ex2_approx(m_used - m_blk);  // m_used = -inf -> 0
          m_used = m_blk;
        }
        const float m_use = (m_used == -INFINITY) ? 0.f : m_used;  // nothing valid so far: p = 0, no NaN
        l_part *= alpha;
        float sum = 0.f;
        uint32_t pk[SC / 2];  // this thread's probabilities, bf16
        // Optional MUFU token of the sub-partition (B200VIT_ATTN_TOKEN=1).  Its two softmax warps (one per query tile)
        // drift into lockstep: both in the exp2 pass at once, sharing the one MUFU, then both in their MUFU-free phases with
        // the unit idle (ncu: 47 % of their samples on MUFU.EX2, the pipe 45 % busy).  Alternating them is worth 5 % stand-
        // alone and 1.5 % inside the tower on most boxes, but on one box it cost 14 % (strict alternation amplifies any
        // stall of either tile), so it is off by default.
        if (mufu_token && (t == 1 || g > 0)) named_barrier(tok_mine, 64);
#pragma unroll
        for (int c = 0; c < SC; c += 32) {
          if (all_valid) {
            // packed fp32x2 FFMA / FADD: two lanes per issue slot for everything except the MUFU itself
            const uint64_t sc2 = f32x2_pack(scale_log2, scale_log2), nm2 = f32x2_pack(-m_use, -m_use);
            uint64_t s2a = 0ull, s2b = 0ull;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float x0, x1, x2, x3;
              f32x2_unpack(f32x2_fma(f32x2_pack_bits(sv[c + i], sv[c + i + 1]), sc2, nm2), x0, x1);
              f32x2_unpack(f32x2_fma(f32x2_pack_bits(sv[c + i + 2], sv[c + i + 3]), sc2, nm2), x2, x3);
              const float p0 = ex2_approx(x0), p1 = ex2_approx(x1), p2 = ex2_approx(x2), p3 = ex2_approx(x3);
              s2a = f32x2_add(s2a, f32x2_pack(p0, p1));
              s2b = f32x2_add(s2b, f32x2_pack(p2, p3));
              pk[(c + i) >> 1] = pack_bf16x2(p0, p1);
              pk[((c + i) >> 1) + 1] = pack_bf16x2(p2, p3);
            }
            float sa, sb, sc, sd;
            f32x2_unpack(s2a, sa, sb);
            f32x2_unpack(s2b, sc, sd);
            sum += (sa + sb) + (sc + sd);
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float p0 = ex2_approx(__uint_as_float(sv[c + i]) * scale_log2 - m_use);
              float p1 = ex2_approx(__uint_as_float(sv[c + i + 1]) * scale_log2 - m_use);
              p0 = (c + i >= lo && c + i < hi) ? p0 : 0.f;
              p1 = (c + i + 1 >= lo && c + i + 1 < hi) ? p1 : 0.f;
              sum += p0 + p1;
              pk[(c + i) >> 1] = pack_bf16x2(p0, p1);
            }
          }
        }
        if (mufu_token) asm volatile("bar.arrive %0, 64;" ::"r"(tok_other) : "memory");  // hand the token over
        TSTAMP2(4);
        if (j >= 1) {
          // every probability of the block is computed: only now is P.V(g-1) needed (P may be overwritten, O is at rest)
          mbar_wait(&o_full[t], (g - 1) & 1);
          TSTAMP2(7);
          if (__any_sync(0xffffffffu, grow)) {  // rare after the first block: O(row) *= 2^(m_old - m_new) in TMEM
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < OC; c0 += 40) {  // 40 columns (32 + 8) per trip
              uint32_t v[32], v8[8];
              tmem_ld32(to + c0, v);
              tmem_ld8(to + c0 + 32, v8);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
#pragma unroll
              for (int i = 0; i < 8; ++i) v8[i] = __float_as_uint(__uint_as_float(v8[i]) * alpha);
              tmem_st32(to + c0, v);
              tmem_st8(to + c0 + 32, v8);
            }
            tmem_st_wait();
#ifdef B200_ATTN_TIMING
            ++n_rescale;
#endif
          }
          TSTAMP2(6);
        } else if (g > 0) {
          // first block of a later item: the previous item's output store (staged in this warp's rows of the P region)
          // must have been read out of shared memory
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
        }
#pragma unroll
        for (int q = 0; q < SC / 8; ++q)  // 16-byte chunk q of this thread's columns: sub-tile q / 8, chunk q % 8
          st_shared_v4(swz128(pbuf + (q >> 3) * T64_BYTES, r, q & 7), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        l_part += sum;
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_full[t]);
        TSTAMP2(5);
#ifdef B200_ATTN_TIMING
        ++n_blocks;
#endif
      }
      // the finished O of this thread's columns
      mbar_wait(&o_full[t], (g - 1) & 1);
      tc_fence_after();
      float o[OC];
#pragma unroll
      for (int c0 = 0; c0 < OC; c0 += 40) {
        uint32_t v[32], v8[8];
        tmem_ld32(to + c0, v);
        tmem_ld8(to + c0 + 32, v8);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c0 + i] = __uint_as_float(v[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[c0 + 32 + i] = __uint_as_float(v8[i]);
      }
      TSTAMP2(8);
      // stage O (bf16) in this warp's own rows of the idle P buffer
      const float inv = (l_part > 0.f) ? 1.f / l_part : 0.f;
#pragma unroll
      for (int c = 0; c < OC; c += 8) {
        const int col = c;
        const uint32_t w0 = pack_bf16x2(o[c] * inv, o[c + 1] * inv), w1 = pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv);
        const uint32_t w2 = pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv), w3 = pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv);
        if (col < 64) st_shared_v4(swz128(stg64, lane, col >> 3), w0, w1, w2, w3);
        else st_shared_v4(stg16 + lane * 32 + (col - 64) * 2, w0, w1, w2, w3);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const int row0 = tile.q_row0 + t * QT + quad * 32;
        tma_store_2d(&to64, stg64, head * HD, row0);  // rows >= m_rows are clipped by the tensor map
        tma_store_2d(&to16, stg16, head * HD + 64, row0);
        bulk_commit();
      }
      TSTAMP2(0);
    }  // items of this CTA
    // tile 1's last hand-over has no taker: tile 0 absorbs it so that no barrier is left half-arrived
    if (mufu_token && t == 0 && g > 0) named_barrier(tok_mine, 64);
    if (lane == 0) bulk_wait<0>();  // output stores have landed before the CTA exits
#ifdef B200_ATTN_TIMING
    if (trec) printf("full attention softmax warp %d (%d blocks, %d rescales): other+store %lld | wait_s %lld | ld+max %lld | - %lld | exps %lld | wait_pv %lld | rescale %lld | store+arrive %lld | final O %lld\n",
                     warp, n_blocks, n_rescale, tstamp[0], tstamp[1], tstamp[2], tstamp[3], tstamp[4], tstamp[7], tstamp[6], tstamp[5], tstamp[8]);
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int QTILES, int NKV, int NQ, bool SHARED_KV>
int launch_variant(const AttnPrepared& g, const AttnTile* d_tiles, int n_tiles, const int2* bd, int m_rows, int heads, int hpc,
                   float scale_log2, cudaStream_t stream) {
  using L = AttnCfg<QTILES, NKV, NQ, SHARED_KV>;
  auto kern = attn_tc_kernel<QTILES, NKV, NQ, SHARED_KV>;
  static DeviceOnce attr;
  if (attr.need()) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES));
    attr.mark();
  }
  B200_CUDA_OK(launch_kernel(kern, dim3(n_tiles, heads / hpc), dim3(L::THREADS), L::BYTES, stream, 1, g.tm64, g.tm16, g.to64,
                             g.to16, d_tiles, bd, m_rows, heads, hpc, scale_log2));
  return 0;
}

}  // namespace

// query tiles of `rows_per_tile` rows + per-row segment bounds from a cu_seqlens list
void build_attn_tiles(const std::vector<int32_t>& cu, int m_rows, int rows_per_tile, std::vector<AttnTile>& tiles,
                      std::vector<int32_t>& bounds) {
  bounds.assign(static_cast<size_t>(m_rows) * 2, 0);
  for (size_t s = 0; s + 1 < cu.size(); ++s)
    for (int r = cu[s]; r < cu[s + 1] && r < m_rows; ++r) {
      bounds[2 * r] = cu[s];
      bounds[2 * r + 1] = cu[s + 1];
    }
  tiles.clear();
  for (int q0 = 0; q0 < m_rows; q0 += rows_per_tile) {
    int lo = INT_MAX, hi = 0;
    for (int r = q0; r < q0 + rows_per_tile && r < m_rows; ++r) {
      if (bounds[2 * r + 1] > bounds[2 * r]) {
        lo = bounds[2 * r] < lo ? bounds[2 * r] : lo;
        hi = bounds[2 * r + 1] > hi ? bounds[2 * r + 1] : hi;
      }
    }
    if (hi <= lo) continue;  // no row of this tile belongs to a segment
    tiles.push_back(AttnTile{q0, lo, (hi - lo + KVB - 1) / KVB, 0});
  }
}

// rows_per_tile = 128 (window layers) or 256 (full layers), matching how `d_tiles` was built
int launch_attention_tc(const void* qkv, void* out, const AttnTile* d_tiles, int n_tiles, int rows_per_tile, int max_blocks,
                        const int32_t* d_bounds, int m_rows, int heads, cudaStream_t stream, AttnPrepared* cache,
                        int32_t* work_counter) {
  if (n_tiles <= 0) return 0;
  AttnPrepared local;
  AttnPrepared& g = cache ? *cache : local;
  const int D = heads * HD;
  if (!(g.valid && g.qkv == qkv && g.out == out && g.m_rows == m_rows && g.heads == heads)) {
    int rc = make_tmap_2d(&g.tm64, qkv, m_rows, 3 * D, 3 * D, 2, 128, 64, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&g.tm16, qkv, m_rows, 3 * D, 3 * D, 2, 128, 16, 32);
    if (rc) return rc;
    rc = make_tmap_2d(&g.to64, out, m_rows, D, D, 2, 32, 64, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&g.to16, out, m_rows, D, D, 2, 32, 16, 0);
    if (rc) return rc;
    g.qkv = qkv, g.out = out, g.m_rows = m_rows, g.heads = heads, g.valid = true;
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  const int2* bd = reinterpret_cast<const int2*>(d_bounds);
  if (rows_per_tile == 256) {
    static int generic = -1, token = -1;
    if (generic < 0) {
      const char* e = getenv("B200VIT_ATTN_FULL_GENERIC");  // debugging: the generic kernel's two-tile variant
      generic = (e != nullptr && e[0] == '1') ? 1 : 0;
      e = getenv("B200VIT_ATTN_TOKEN");
      token = (e != nullptr && e[0] == '1') ? 1 : 0;  // off by default, see the exp2 pass
    }
    if (generic) return launch_variant<2, 2, 1, true>(g, d_tiles, n_tiles, bd, m_rows, heads, 1, scale_log2, stream);
    using L = AttnCfg<2, 2, 1, true>;
    static DeviceOnce attr;
    if (attr.need()) {
      B200_CUDA_OK(cudaFuncSetAttribute(attn_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES));
      attr.mark();
    }
    const int n_items = n_tiles * heads;
    int grid = n_items < device_sm_count() ? n_items : device_sm_count();  // persistent: one CTA per SM
    static int one_item = -1;
    if (one_item < 0) {
      const char* e = getenv("B200VIT_ATTN_ONE_ITEM");  // experiment: one CTA per item (hardware-balanced)
      one_item = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    if (one_item) grid = n_items;
    static int static_walk = -1;
    if (static_walk < 0) {
      const char* e = getenv("B200VIT_ATTN_STATIC");  // experiment: static stride instead of the work counter
      static_walk = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    if (static_walk) work_counter = nullptr;
    B200_CUDA_OK(launch_kernel(attn_full_kernel, dim3(grid), dim3(F_THREADS), L::BYTES, stream, 1, g.tm64, g.tm16, g.to64, g.to16,
                               d_tiles, n_tiles, bd, m_rows, heads, scale_log2, token, work_counter));
    return 0;
  }
  if (rows_per_tile != 128) return fail(B200VIT_EINVAL, "attention: rows_per_tile must be 128 or 256");
  (void)max_blocks;
  if (heads % 2 != 0)  // odd head count: single stream, K/V double-buffered
    return launch_variant<1, 2, 2, true>(g, d_tiles, n_tiles, bd, m_rows, heads, 1, scale_log2, stream);
  // window layers: two heads in flight per CTA; walk as many head pairs per CTA as keep ~one CTA per SM
  int hpc = 2;
  const int sms = device_sm_count();
  for (int c = 8; c >= 4; c >>= 1)
    if (heads % c == 0 && n_tiles * (heads / c) >= (sms * 3) / 4) {
      hpc = c;
      break;
    }
  return launch_variant<2, 1, 1, false>(g, d_tiles, n_tiles, bd, m_rows, heads, hpc, scale_log2, stream);
}

}  // namespace b200
