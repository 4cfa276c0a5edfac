// Frame resize ahead of the overlay: Pillow's bicubic Image.resize on uint8 RGB frames, bit for bit.
//
// Reference call site: qwen_vl_utils.process_vision_info (/root/reference/app.py:296, :417, utils/dataset.py:76), whose
// fetch_image does  smart_resize(...)  then  image.resize((w, h))  -- Pillow's default BICUBIC.  Pillow's resampler
// (src/libImaging/Resample.c) is two separable integer passes: per output coordinate a window of 22-bit fixed-point
// coefficients (precompute_coeffs + normalize_coeffs_8bpc, computed in double exactly as Pillow does),
// horizontal pass into an intermediate uint8 image, then vertical pass, each  clip8((sum + 2^21) >> 22).
// See oracle/resize_ref.py for the restatement that is pinned against PIL.
//
// HBM-bound byte work: 3 B/px in, 3 B/px out plus the intermediate image (written and read once).
#include <algorithm>
#include <cmath>
#include <cstdint>

#include "internal.h"

namespace b200 {
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

// Coefficients (Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole-image box) are computed on the
// device, one thread per output coordinate, so a call enqueues no host->device copies.  bounds: (first, count).
// Double arithmetic with explicit round-to-nearest intrinsics (no FMA contraction) is IEEE-identical to Pillow's C.
__device__ double bicubic_filter_dev(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(a + 2.0, x), a + 3.0), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), a);
  return 0.0;
}

// One warp per output coordinate, lane = tap: every weight is evaluated once; the normalising sum is accumulated in tap
// order by lane 0 exactly like Pillow's loop (floating-point addition order matters for bit-exactness).
__device__ void resize_coeffs_one(int in_size, int out_size, int ksize, int xx, int32_t* __restrict__ bounds,
                                  int32_t* __restrict__ kk) {
  const int lane = threadIdx.x & 31;
  const double scale = static_cast<double>(static_cast<float>(in_size) - 0.0f) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, filterscale);
  const double center = __dadd_rn(0.0, __dmul_rn(xx + 0.5, scale));
  const double ss = 1.0 / filterscale;
  int xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double ww = 0.0;
  for (int x0 = 0; x0 < xmax; x0 += 32) {  // sequential sum over the taps, 32 at a time
    const int x = x0 + lane;
    const double w = x < xmax ? bicubic_filter_dev(__dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss)) : 0.0;
    const int n = min(32, xmax - x0);
    for (int i = 0; i < n; ++i) ww = __dadd_rn(ww, __shfl_sync(0xffffffffu, w, i));
  }
  int32_t* k = kk + static_cast<size_t>(xx) * ksize;
  for (int x = lane; x < ksize; x += 32) {
    int32_t q = 0;
    if (x < xmax) {
      double v = bicubic_filter_dev(__dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss));
      if (ww != 0.0) v = v / ww;
      const double fixed = __dmul_rn(v, static_cast<double>(1 << PRECISION_BITS));
      q = v < 0 ? static_cast<int32_t>(__dadd_rn(-0.5, fixed)) : static_cast<int32_t>(__dadd_rn(0.5, fixed));
    }
    k[x] = q;
  }
  if (lane == 0) {
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
}

// warps [0, w_out) build the horizontal table, warps [w_out, w_out + h_out) the vertical one (either may be skipped)
__global__ void __launch_bounds__(128)
resize_coeffs_kernel(int w_in, int w_out, int ksize_h, int32_t* bounds_h, int32_t* kk_h, int h_in, int h_out, int ksize_v,
                     int32_t* bounds_v, int32_t* kk_v, int need_h, int need_v) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid < w_out) {
    if (need_h) resize_coeffs_one(w_in, w_out, ksize_h, wid, bounds_h, kk_h);
  } else if (wid < w_out + h_out) {
    if (need_v) resize_coeffs_one(h_in, h_out, ksize_v, wid - w_out, bounds_v, kk_v);
  }
}

__device__ __forceinline__ uint8_t clip8(int v) { return static_cast<uint8_t>(min(max(v >> PRECISION_BITS, 0), 255)); }

// Horizontal pass.  One block = 256 consecutive output pixels of RESIZE_H_ROWS consecutive rows: the input spans they
// need (contiguous bytes per row) are staged in shared memory with coalesced 4-byte loads, then every thread runs its
// taps from shared memory, loading each coefficient once for all its rows.
constexpr int RESIZE_H_ROWS = 4;
constexpr int RESIZE_H_SMEM_MAX = 96 * 1024;

__global__ void __launch_bounds__(256)
resize_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int rows_total, int w_in, int w_out,
                const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize, int span_bytes) {
  extern __shared__ __align__(16) uint8_t span[];
  const int64_t row0 = static_cast<int64_t>(blockIdx.y) * RESIZE_H_ROWS;
  const int nrows = static_cast<int>(min(static_cast<int64_t>(RESIZE_H_ROWS), rows_total - row0));
  const int xo0 = blockIdx.x * 256;
  const int xo1 = min(xo0 + 256, w_out) - 1;
  const int first = __ldg(bounds + 2 * xo0);
  const int last = __ldg(bounds + 2 * xo1) + __ldg(bounds + 2 * xo1 + 1);  // exclusive
  const int64_t row_bytes = static_cast<int64_t>(w_in) * 3;
  const uint8_t* src0 = in + row0 * row_bytes;
  // staging needs every row start 4-byte aligned: the base pointer and the row pitch
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | static_cast<uintptr_t>(row_bytes)) & 3) == 0;
  const int64_t b0 = (static_cast<int64_t>(first) * 3) & ~int64_t(3);
  const int nbytes = static_cast<int>(static_cast<int64_t>(last) * 3 - b0);
  const bool staged = span_bytes > 0 && nbytes <= span_bytes && aligned;
  if (staged) {
    // all rows' loads of one word column are issued before any store: the pass is latency-bound otherwise (ncu: a
    // load->store loop keeps one request in flight per warp)
    const int full = nbytes >> 2;
    for (int i = threadIdx.x; i < full; i += 256) {
      uint32_t tmp[RESIZE_H_ROWS];
#pragma unroll
      for (int r = 0; r < RESIZE_H_ROWS; ++r)
        tmp[r] = r < nrows ? __ldg(reinterpret_cast<const uint32_t*>(src0 + r * row_bytes + b0) + i) : 0u;
#pragma unroll
      for (int r = 0; r < RESIZE_H_ROWS; ++r) reinterpret_cast<uint32_t*>(span + r * span_bytes)[i] = tmp[r];
    }
    for (int i = (full << 2) + threadIdx.x; i < nbytes; i += 256)
      for (int r = 0; r < nrows; ++r) span[r * span_bytes + i] = __ldg(src0 + r * row_bytes + b0 + i);
    __syncthreads();
  }
  const int xo = xo0 + threadIdx.x;
  if (xo >= w_out) return;
  const int xmin = __ldg(bounds + 2 * xo), cnt = __ldg(bounds + 2 * xo + 1);
  const int32_t* k = kk + static_cast<size_t>(xo) * ksize;
  int acc[RESIZE_H_ROWS][3];
#pragma unroll
  for (int r = 0; r < RESIZE_H_ROWS; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 1 << (PRECISION_BITS - 1);
  if (staged) {
    const uint8_t* p = span + (static_cast<int64_t>(xmin) * 3 - b0);
    for (int x = 0; x < cnt; ++x) {
      const int c = __ldg(k + x);
#pragma unroll
      for (int r = 0; r < RESIZE_H_ROWS; ++r) {  // rows beyond nrows read stale shared memory and are never stored
        const uint8_t* q = p + r * span_bytes + 3 * x;
        acc[r][0] += static_cast<int>(q[0]) * c;
        acc[r][1] += static_cast<int>(q[1]) * c;
        acc[r][2] += static_cast<int>(q[2]) * c;
      }
    }
  } else {
    for (int x = 0; x < cnt; ++x) {
      const int c = __ldg(k + x);
      for (int r = 0; r < nrows; ++r) {
        const uint8_t* q = src0 + r * row_bytes + static_cast<int64_t>(xmin + x) * 3;
        acc[r][0] += static_cast<int>(__ldg(q)) * c;
        acc[r][1] += static_cast<int>(__ldg(q + 1)) * c;
        acc[r][2] += static_cast<int>(__ldg(q + 2)) * c;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RESIZE_H_ROWS; ++r)
    if (r < nrows) {
      uint8_t* dst = out + ((row0 + r) * w_out + xo) * 3;
      dst[0] = clip8(acc[r][0]), dst[1] = clip8(acc[r][1]), dst[2] = clip8(acc[r][2]);
    }
}

// one thread = four consecutive output bytes of a row (the row pitch w*3 is a multiple of 4 when w is)
__global__ void __launch_bounds__(256)
resize_v_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int frames, int h_in, int h_out, int row_bytes,
                const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize) {
  const int words = row_bytes >> 2;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<int64_t>(frames) * h_out * words) return;
  const int wi = static_cast<int>(gid % words);
  const int yo = static_cast<int>((gid / words) % h_out);
  const int f = static_cast<int>(gid / (static_cast<int64_t>(words) * h_out));
  const int ymin = __ldg(bounds + 2 * yo), cnt = __ldg(bounds + 2 * yo + 1);
  const int32_t* k = kk + static_cast<size_t>(yo) * ksize;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (static_cast<size_t>(f) * h_in + ymin) * row_bytes) + wi;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0, s3 = s0;
  for (int y0 = 0; y0 < cnt; y0 += 4) {  // four taps' loads in flight together (a zero coefficient pads the tail)
    int c[4];
    uint32_t v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool on = y0 + u < cnt;
      c[u] = on ? __ldg(k + y0 + u) : 0;
      v[u] = on ? __ldg(src + static_cast<size_t>(y0 + u) * words) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s0 += static_cast<int>(v[u] & 0xffu) * c[u];
      s1 += static_cast<int>((v[u] >> 8) & 0xffu) * c[u];
      s2 += static_cast<int>((v[u] >> 16) & 0xffu) * c[u];
      s3 += static_cast<int>(v[u] >> 24) * c[u];
    }
  }
  reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(f) * h_out + yo) * row_bytes)[wi] =
      clip8(s0) | (clip8(s1) << 8) | (clip8(s2) << 16) | (static_cast<uint32_t>(clip8(s3)) << 24);
}

size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

struct ResizeLayout {
  size_t tmp, kh, bh, kv, bv, total;
  int ksize_h, ksize_v;
};
int ksize_for(int in_size, int out_size) {
  double fs = static_cast<double>(static_cast<float>(in_size) - 0.0f) / out_size;
  if (fs < 1.0) fs = 1.0;
  return static_cast<int>(std::ceil(2.0 * fs)) * 2 + 1;
}
ResizeLayout resize_layout(int t, int h_in, int w_in, int h_out, int w_out) {
  ResizeLayout l;
  l.ksize_h = ksize_for(w_in, w_out);
  l.ksize_v = ksize_for(h_in, h_out);
  size_t off = 0;
  l.tmp = off, off += align256(static_cast<size_t>(t) * h_in * w_out * 3);
  l.kh = off, off += align256(static_cast<size_t>(w_out) * l.ksize_h * 4);
  l.bh = off, off += align256(static_cast<size_t>(w_out) * 8);
  l.kv = off, off += align256(static_cast<size_t>(h_out) * l.ksize_v * 4);
  l.bv = off, off += align256(static_cast<size_t>(h_out) * 8);
  l.total = off;
  return l;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" size_t b200vit_resize_workspace_bytes(int32_t t, int32_t h_in, int32_t w_in, int32_t h_out, int32_t w_out) {
  if (t <= 0 || h_in <= 0 || w_in <= 0 || h_out <= 0 || w_out <= 0) return 0;
  return resize_layout(t, h_in, w_in, h_out, w_out).total;
}

extern "C" int b200vit_resize_bicubic(const uint8_t* d_in, int32_t t, int32_t h_in, int32_t w_in, uint8_t* d_out, int32_t h_out,
                                      int32_t w_out, void* d_workspace, size_t workspace_bytes, b200vit_stream stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = check_arch();
  if (rc) return rc;
  if (d_in == nullptr || d_out == nullptr || t <= 0 || h_in <= 0 || w_in <= 0 || h_out <= 0 || w_out <= 0)
    return fail(B200VIT_EINVAL, "resize: empty clip or null pointer");
  if (h_in > 16384 || w_in > 16384 || h_out > 16384 || w_out > 16384) return fail(B200VIT_EINVAL, "resize: side exceeds 16384");
  if (w_out % 4) return fail(B200VIT_EINVAL, "resize: output width must be a multiple of 4 (it is a multiple of 28 on this path)");
  const ResizeLayout lay = resize_layout(t, h_in, w_in, h_out, w_out);
  if (d_workspace == nullptr || workspace_bytes < lay.total) return fail(B200VIT_EINVAL, "resize: workspace too small");
  if ((reinterpret_cast<uintptr_t>(d_workspace) & 255) || (reinterpret_cast<uintptr_t>(d_out) & 3))
    return fail(B200VIT_EALIGN, "resize: workspace must be 256-byte and output 4-byte aligned");
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const bool need_h = w_out != w_in, need_v = h_out != h_in;
  if (!need_h && !need_v) {
    B200_CUDA_OK(cudaMemcpyAsync(d_out, d_in, static_cast<size_t>(t) * h_in * w_in * 3, cudaMemcpyDeviceToDevice, stream));
    return 0;
  }
  const uint8_t* src = d_in;
  int32_t* kh = reinterpret_cast<int32_t*>(ws + lay.kh);
  int32_t* bh = reinterpret_cast<int32_t*>(ws + lay.bh);
  int32_t* kv = reinterpret_cast<int32_t*>(ws + lay.kv);
  int32_t* bv = reinterpret_cast<int32_t*>(ws + lay.bv);
  {
    const int warps = w_out + h_out;
    resize_coeffs_kernel<<<(warps * 32 + 127) / 128, 128, 0, stream>>>(w_in, w_out, lay.ksize_h, bh, kh, h_in, h_out, lay.ksize_v,
                                                                       bv, kv, need_h ? 1 : 0, need_v ? 1 : 0);
  }
  if (need_h) {
    uint8_t* dst = need_v ? ws + lay.tmp : d_out;
    static DeviceOnce attr;
    if (attr.need()) {
      B200_CUDA_OK(cudaFuncSetAttribute(resize_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RESIZE_H_SMEM_MAX));
      attr.mark();
    }
    // input bytes 256 consecutive outputs can touch: 256 steps of `scale` plus the filter support on both sides
    const double scale = static_cast<double>(w_in) / w_out;
    const double span_px = 256.0 * scale + 2.0 * (2.0 * (scale < 1.0 ? 1.0 : scale)) + 4.0;
    int span_bytes = (static_cast<int>(span_px) * 3 + 8 + 15) & ~15;
    if (span_bytes * RESIZE_H_ROWS > RESIZE_H_SMEM_MAX) span_bytes = 0;  // enormous down-scale: taps straight from global memory
    const int64_t rows_total = static_cast<int64_t>(t) * h_in;
    const int64_t row_blocks = (rows_total + RESIZE_H_ROWS - 1) / RESIZE_H_ROWS;
    // blockIdx.y is limited to 65535: fold row blocks into launches of at most that many
    for (int64_t rb0 = 0; rb0 < row_blocks; rb0 += 65535) {
      const int nb = static_cast<int>(std::min<int64_t>(65535, row_blocks - rb0));
      const int64_t r0 = rb0 * RESIZE_H_ROWS;
      resize_h_kernel<<<dim3((w_out + 255) / 256, nb), 256, span_bytes * RESIZE_H_ROWS, stream>>>(
          src + r0 * w_in * 3, dst + r0 * w_out * 3, static_cast<int>(std::min<int64_t>(rows_total - r0, 65535LL * RESIZE_H_ROWS)),
          w_in, w_out, bh, kh, lay.ksize_h, span_bytes);
    }
    src = dst;
  }
  if (need_v) {
    if (reinterpret_cast<uintptr_t>(src) & 3) return fail(B200VIT_EALIGN, "resize: input must be 4-byte aligned");
    const int64_t n = static_cast<int64_t>(t) * h_out * (w_out * 3 / 4);
    resize_v_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(src, d_out, t, h_in, h_out, w_out * 3, bv, kv,
                                                                                  lay.ksize_v);
  }
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}
