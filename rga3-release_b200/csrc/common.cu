// Error plumbing, device queries and TMA tensor-map creation shared by all kernels.
#include <mutex>

#include "internal.h"

namespace b200 {

static thread_local std::string g_err;

void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + what;
  return B200VIT_ECUDA;
}
const char* last_error_cstr() { return g_err.c_str(); }

int device_sm_count() {
  static std::atomic<int> sms[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int v = sms[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    sms[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

int check_arch() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(B200VIT_ECUDA, "no CUDA device");
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return fail(B200VIT_EARCH, "libb200vit is built for sm_100a only; device is not compute capability 10.x");
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    // resolved through the runtime so the library does not link libcuda.so
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(B200VIT_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  if (reinterpret_cast<uintptr_t>(base) & 15) return fail(B200VIT_EALIGN, "TMA base pointer must be 16-byte aligned");
  if ((cols * 2) % 16) return fail(B200VIT_EALIGN, "TMA row pitch must be a multiple of 16 bytes");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200VIT_ECUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string(int(r)) + ")");
  return 0;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t pitch_elems, int elem_bytes,
                 int box_rows, int box_cols, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(B200VIT_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  if (reinterpret_cast<uintptr_t>(base) & 15) return fail(B200VIT_EALIGN, "TMA base pointer must be 16-byte aligned");
  if ((pitch_elems * elem_bytes) % 16) return fail(B200VIT_EALIGN, "TMA row pitch must be a multiple of 16 bytes");
  if (swizzle_bytes != 0 && swizzle_bytes != 32 && swizzle_bytes != 64 && swizzle_bytes != 128) return fail(B200VIT_EINVAL, "unsupported TMA swizzle");
  if (swizzle_bytes && box_cols * elem_bytes != swizzle_bytes) return fail(B200VIT_EINVAL, "swizzled TMA box must span the swizzle width");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(pitch_elems) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200VIT_ECUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string(int(r)) + ")");
  return 0;
}

}  // namespace b200

extern "C" int b200vit_version(void) { return B200VIT_VERSION; }
namespace b200 { const char* last_error_cstr(); }
extern "C" const char* b200vit_last_error(void) { return b200::last_error_cstr(); }
