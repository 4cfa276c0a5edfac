// Internal (C++) interfaces between the translation units of libb200vit.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/b200vit.h"

namespace b200 {

// thread-local error string behind b200vit_last_error()
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define B200_CUDA_OK(expr)                                  \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return ::b200::cuda_fail(_e, #expr); \
  } while (0)

// "Once per device" guard for per-device state (function attributes, __device__ symbol uploads): a process may
// drive several GPUs, and cudaFuncSetAttribute / cudaMemcpyToSymbol only affect the current one.  Redoing the set-up
// from two threads at once is harmless (idempotent), so no lock.
struct DeviceOnce {
  std::atomic<unsigned long long> done{0};
  int dev = 0;
  bool need() {
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    return dev < 0 || dev >= 64 || ((done.load(std::memory_order_acquire) >> dev) & 1ull) == 0;
  }
  void mark() {
    if (dev >= 0 && dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
  }
};

int device_sm_count();  // of the current device
int check_arch();  // 0 if the current device is sm_100, else B200VIT_EARCH

// 2-D bf16 row-major [rows, cols] tensor map, box = [box_rows, 64 cols], SWIZZLE_128B,
// out-of-bounds elements read as zero.
int make_tmap_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int box_rows);
// General 2-D row-major map: elem_bytes 2 (bf16) or 4 (fp32), row pitch `pitch_elems`, box = [box_rows, box_cols],
// swizzle_bytes = 0 (none), 32, 64 or 128 (box_cols * elem_bytes must equal the swizzle width).
int make_tmap_2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t pitch_elems, int elem_bytes,
                 int box_rows, int box_cols, int swizzle_bytes);

// kernels (all enqueue on `stream`, no host sync)
// Host-side memo of one GEMM call site: tensor maps and launch geometry are rebuilt only when the
// arguments change (cuTensorMapEncodeTiled x3 per launch made the host the bottleneck otherwise).
struct GemmPrepared {
  bool valid = false;
  b200vit_gemm_args key;
  CUtensorMap ta, tb, to, taux;
  int grid = 0, stream_k = 0;
};
int launch_gemm(const b200vit_gemm_args& a, cudaStream_t stream, GemmPrepared* cache = nullptr);
int launch_rmsnorm(const float* x, const float* w, void* out_bf16, int rows, int dim, float eps, cudaStream_t stream);
int launch_cast_bf16(const void* in, int in_dtype, void* out, int64_t n, cudaStream_t stream);

// tcgen05 attention (attention_tc.cu): one CTA = 128 query rows of one head
struct AttnTile {
  int32_t q_row0, kv_row0, n_kv_blocks, pad;
};
struct AttnPrepared {  // host memo of the two tensor maps over the qkv buffer
  bool valid = false;
  const void* qkv = nullptr;
  const void* out = nullptr;
  int m_rows = 0, heads = 0;
  CUtensorMap tm64, tm16;  // loads from the qkv buffer
  CUtensorMap to64, to16;  // stores to the attention output
};
void build_attn_tiles(const std::vector<int32_t>& cu, int m_rows, int rows_per_tile, std::vector<AttnTile>& tiles,
                      std::vector<int32_t>& bounds);
int launch_attention_tc(const void* qkv, void* out, const AttnTile* d_tiles, int n_tiles, int rows_per_tile, int max_blocks,
                        const int32_t* d_bounds, int m_rows, int heads, cudaStream_t stream, AttnPrepared* cache,
                        int32_t* work_counter = nullptr);  // full layers: a zeroed device int -> dynamic work items

struct OverlayDev;  // device-side overlay description (overlay.cu)
int launch_overlay_patchify(const b200vit_frames& fr, const b200vit_overlay* ov, int patch, int tps, int merge,
                            void* out_bf16, uint8_t* out_u8, cudaStream_t stream);

}  // namespace b200
