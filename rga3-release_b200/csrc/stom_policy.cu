// STOM placement policy on the device: tracker outputs -> one frame op per frame, no host round trip.
//
// Reference: /root/reference/model/STOM.py
//   propagate_in_video :72-141   which frame gets which overlay
//   flow filter        :104-131  MAD filter on the flows of the visible tracks, mean flow (numpy float32)
//   warp_point         :163-203  visible-track mask -> cv2 closing (ellipse k = min(h,w)/15) -> moments centroid
// The arithmetic is reproduced bit for bit: every float32 operation numpy performs is one IEEE float32 operation
// here (no FMA contraction: __fmul_rn/__fadd_rn/...), np.median = sort + middle, np.mean = numpy's pairwise summation
// order + one division, cv2.moments = exact integer sums divided in double.  OpenCV's structuring element and
// dilate/erode conventions are restated in oracle/stom_policy_ref.py and checked against cv2 there.
//
// Work is tiny (T frames x N <= 16384 points, T masks of h x w bytes); the point is the missing D2H -> numpy/cv2 -> H2D
// hop between the tracker and the overlay kernel, so one CTA per frame is enough.
#include <cmath>
#include <cstdint>
#include <cstring>

#include "internal.h"
#include "launch.cuh"

namespace b200 {
namespace {

static_assert(sizeof(b200vit_frame_op) == 36, "b200vit_frame_op is read as 9 x 32-bit words by overlay.cu");

constexpr int MAX_POINTS = 16384;
constexpr int POLICY_THREADS = 1024;
constexpr int MAX_SE = 128;  // structuring element rows: min(h,w)/15

struct SeRows {
  int32_t k, anchor;
  int16_t j1[MAX_SE], j2[MAX_SE];
};

__device__ __forceinline__ void write_op(b200vit_frame_op* dst, int mode, int sx, int sy, int zx, int zy, int cx, int cy,
                                         int r, uint32_t rgba) {
  int32_t* d = reinterpret_cast<int32_t*>(dst);
  d[0] = mode, d[1] = sx, d[2] = sy, d[3] = zx, d[4] = zy, d[5] = cx, d[6] = cy, d[7] = r;
  d[8] = static_cast<int32_t>(rgba);
}

// in-place ascending bitonic sort of s[0, p2) (p2 a power of two), all threads of the block
__device__ void bitonic_sort(float* s, int p2) {
  for (int k = 2; k <= p2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < p2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = s[i], b = s[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) s[i] = b, s[ixj] = a;
        }
      }
      __syncthreads();
    }
  }
}

// np.median of m finite float32 values held (unsorted) in src[0, m); sorts a padded copy in `sorted`
__device__ float median_f32(const float* src, int m, float* sorted) {
  int p2 = 1;
  while (p2 < m) p2 <<= 1;
  for (int i = threadIdx.x; i < p2; i += blockDim.x) sorted[i] = i < m ? src[i] : INFINITY;
  __syncthreads();
  bitonic_sort(sorted, p2);
  const float med = (m & 1) ? sorted[m >> 1] : __fdiv_rn(__fadd_rn(sorted[(m >> 1) - 1], sorted[m >> 1]), 2.0f);
  __syncthreads();
  return med;
}

// numpy's float32 pairwise add.reduce (loops_utils.h pairwise_sum) over a[0, n) with element stride `st`, one thread
__device__ float pairwise_sum_f32(const float* a, int n, int st) {
  // explicit stack of pending [offset, length) ranges, processed left to right; partial sums are combined in the
  // same tree order as the recursion  sum(a[:n2]) + sum(a[n2:])
  struct Item { int off, len, state; float left; };
  Item stack[24];
  int sp = 0;
  stack[sp++] = {0, n, 0, 0.f};
  float ret = 0.f;
  while (sp > 0) {
    Item& it = stack[sp - 1];
    if (it.len <= 128) {
      const float* p = a + static_cast<size_t>(it.off) * st;
      float res;
      if (it.len < 8) {
        res = -0.0f;
        for (int i = 0; i < it.len; ++i) res = __fadd_rn(res, p[static_cast<size_t>(i) * st]);
      } else {
        float r[8];
        for (int k = 0; k < 8; ++k) r[k] = p[static_cast<size_t>(k) * st];
        int i = 8;
        for (; i < it.len - (it.len % 8); i += 8)
          for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], p[static_cast<size_t>(i + k) * st]);
        res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < it.len; ++i) res = __fadd_rn(res, p[static_cast<size_t>(i) * st]);
      }
      ret = res;
      --sp;
      continue;
    }
    int n2 = it.len / 2;
    n2 -= n2 % 8;
    if (it.state == 0) {         // descend into the left half
      it.state = 1;
      stack[sp++] = {it.off, n2, 0, 0.f};
    } else if (it.state == 1) {  // left done: remember it, descend into the right half
      it.left = ret;
      it.state = 2;
      stack[sp++] = {it.off + n2, it.len - n2, 0, 0.f};
    } else {                     // both done
      ret = __fadd_rn(it.left, ret);
      --sp;
    }
  }
  return ret;
}

// Integer form of STOM.warp's int(x + flow) along one axis of length n (overlay.py::shift_from_flow).
__device__ void shift_from_flow(float flow32, int n, int& shift, int& zero_extra) {
  const double f = static_cast<double>(flow32);  // int64 + float32 promotes to float64 in the reference
  if (!(fabs(f) < 1048576.0)) {                  // far outside any frame (the reference's int(inf) raises)
    shift = f > 0 ? (1 << 20) : -(1 << 20);
    zero_extra = 0;
    return;
  }
  const double fl = floor(f);
  shift = static_cast<int>(fl);
  // the source x = floor(-f) lands in (-1, 0) and truncates to destination 0 as well
  zero_extra = (f < 0.0 && f != fl && floor(-f) < static_cast<double>(n)) ? 1 : 0;
}

// ------------------------------------------------------------------ flow policy (non-mask shapes), one CTA per frame
// scratch per frame: fx[N], fy[N] (visible flows, original order), kx[N], ky[N] (kept flows), mag[N], dev[N]
__global__ void __launch_bounds__(POLICY_THREADS)
stom_flow_kernel(const float* __restrict__ tracks, const uint8_t* __restrict__ vis, int n, int key_idx, int h, int w,
                 float* __restrict__ scratch, b200vit_frame_op* __restrict__ ops) {
  extern __shared__ float sorted[];  // next_pow2(n) floats
  __shared__ int scan_base, warp_tot[POLICY_THREADS / 32];
  __shared__ float sh_mean[2];
  const int f = blockIdx.x;
  if (f == key_idx) {
    if (threadIdx.x == 0) write_op(ops + f, B200VIT_FRAME_LAYER, 0, 0, 0, 0, 0, 0, 0, 0u);
    return;
  }
  float* fx = scratch + static_cast<size_t>(f) * 6 * n;
  float* fy = fx + n;
  float* kx = fy + n;
  float* ky = kx + n;
  float* mag = ky + n;
  float* dev = mag + n;
  const float* trk = tracks + static_cast<size_t>(f) * n * 2;
  const float* key = tracks + static_cast<size_t>(key_idx) * n * 2;
  const uint8_t* v = vis + static_cast<size_t>(f) * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // order-preserving compaction: out position = number of selected elements before i
  auto compact = [&](auto selected, auto emit) -> int {
    if (threadIdx.x == 0) scan_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
      const int i = i0 + threadIdx.x;
      const bool sel = i < n && selected(i);
      const unsigned ballot = __ballot_sync(0xffffffffu, sel);
      if (lane == 0) warp_tot[warp] = __popc(ballot);
      __syncthreads();
      int before = scan_base;
      for (int wi = 0; wi < warp; ++wi) before += warp_tot[wi];
      if (sel) emit(i, before + __popc(ballot & ((1u << lane) - 1)));
      __syncthreads();
      if (threadIdx.x == 0) {
        int tot = 0;
        for (int wi = 0; wi < POLICY_THREADS / 32; ++wi) tot += warp_tot[wi];
        scan_base += tot;
      }
      __syncthreads();
    }
    return scan_base;
  };

  // flows of the visible tracks (:104-106) and their magnitudes (np.linalg.norm(axis=1), :111)
  const int m = compact([&](int i) { return v[i] != 0; },
                        [&](int i, int pos) {
                          const float dx = __fsub_rn(trk[2 * i], key[2 * i]), dy = __fsub_rn(trk[2 * i + 1], key[2 * i + 1]);
                          fx[pos] = dx, fy[pos] = dy;
                          mag[pos] = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
                        });
  if (m == 0) {  // :108-110
    if (threadIdx.x == 0) write_op(ops + f, B200VIT_FRAME_NONE, 0, 0, 0, 0, 0, 0, 0, 0u);
    return;
  }
  __syncthreads();
  // np.median propagates NaN; with a NaN median every comparison below is false and nothing is kept
  int has_nan = 0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) has_nan |= isnan(mag[i]) ? 1 : 0;
  has_nan = __syncthreads_or(has_nan);
  int kept = 0;
  if (!has_nan) {
    const float med = median_f32(mag, m, sorted);                         // :112
    for (int i = threadIdx.x; i < m; i += blockDim.x) dev[i] = fabsf(__fsub_rn(mag[i], med));
    __syncthreads();
    const float mad = median_f32(dev, m, sorted);                         // :113
    const float thr = __fmul_rn(3.0f, mad);                               // :114
    const float lo = __fsub_rn(med, thr), hi = __fadd_rn(med, thr);
    kept = compact([&](int i) { return i < m && mag[i] >= lo && mag[i] <= hi; },   // :115-118 (n >= m: extra i are skipped)
                   [&](int i, int pos) { kx[pos] = fx[i], ky[pos] = fy[i]; });
  }
  if (kept < n / 2) {  // :122-124
    if (threadIdx.x == 0) write_op(ops + f, B200VIT_FRAME_NONE, 0, 0, 0, 0, 0, 0, 0, 0u);
    return;
  }
  __syncthreads();
  // np.mean over the kept flows (:126-127): pairwise float32 sum, one float32 division; x and y on two warps
  if (threadIdx.x == 0 || threadIdx.x == 32) {
    const float* col = threadIdx.x == 0 ? kx : ky;
    sh_mean[threadIdx.x >> 5] = kept > 0 ? __fdiv_rn(pairwise_sum_f32(col, kept, 1), static_cast<float>(kept)) : 0.0f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float mx = sh_mean[0], my = sh_mean[1];
    if (isnan(mx) || isnan(my)) {  // :129-131
      write_op(ops + f, B200VIT_FRAME_NONE, 0, 0, 0, 0, 0, 0, 0, 0u);
    } else {
      int sx, zx, sy, zy;
      shift_from_flow(mx, w, sx, zx);
      shift_from_flow(my, h, sy, zy);
      write_op(ops + f, B200VIT_FRAME_LAYER, sx, sy, zx, zy, 0, 0, 0, 0u);
    }
  }
}

// ------------------------------------------------------------------ point policy (mask shapes)
// frame_state[f]: 0 = proceed, 1 = leave the frame untouched (too few visible tracks :165-166, or a visible
// non-finite coordinate, where the reference raises and its caller keeps the frame :93-100)
constexpr int POINTS_PER_BLOCK = 32;

__global__ void __launch_bounds__(256)
stom_points_kernel(const float* __restrict__ tracks, const uint8_t* __restrict__ vis, int n, int key_idx, int h, int w,
                   SeRows se, uint8_t* __restrict__ dil, int32_t* __restrict__ frame_state) {
  const int f = blockIdx.y;
  if (f == key_idx) return;
  const float* trk = tracks + static_cast<size_t>(f) * n * 2;
  const uint8_t* v = vis + static_cast<size_t>(f) * n;
  // every block of a frame re-derives the frame's verdict (n <= 16384 bytes to scan) instead of waiting for one
  int cnt = 0, bad = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (v[i]) {
      ++cnt;
      bad |= (!isfinite(trk[2 * i]) || !isfinite(trk[2 * i + 1])) ? 1 : 0;
    }
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  if (cnt) atomicAdd(&s_cnt, cnt);
  bad = __syncthreads_or(bad);  // also orders the atomics before the read below
  const bool skip = s_cnt < n / 2 || bad;
  if (blockIdx.x == 0 && threadIdx.x == 0) frame_state[f] = skip ? 1 : 0;
  if (skip) return;
  // dilation of the point mask, stamped point by point: a point at (row, col) sets dst(row - (i - a), col - (j - a))
  // for every element (i, j) of the structuring element.  All writers store 1, so overlaps need no ordering.
  uint8_t* d = dil + static_cast<size_t>(f) * h * w;
  const int p0 = blockIdx.x * POINTS_PER_BLOCK;
  const int work = min(POINTS_PER_BLOCK, n - p0) * se.k;
  for (int idx = threadIdx.x; idx < work; idx += blockDim.x) {
    const int i = p0 + idx / se.k, si = idx % se.k;
    if (!v[i] || se.j1[si] >= se.j2[si]) continue;
    const float px = trk[2 * i + 1], py = trk[2 * i];  // the reference's x is the ROW (:181-185)
    if (fabsf(px) >= 1.0e9f || fabsf(py) >= 1.0e9f) continue;
    const int row = static_cast<int>(truncf(px)), col = static_cast<int>(truncf(py));  // int() truncates toward zero
    if (row < 0 || row >= h || col < 0 || col >= w) continue;
    const int yy = row - (si - se.anchor);
    if (yy < 0 || yy >= h) continue;
    const int x0 = max(col - (se.j2[si] - 1) + se.anchor, 0), x1 = min(col - se.j1[si] + se.anchor + 1, w);
    for (int x = x0; x < x1; ++x) d[static_cast<size_t>(yy) * w + x] = 1;
  }
}

// per-row inclusive prefix counts of the dilated mask: pre[f][y][x+1] = #set in row y, columns [0, x]
__global__ void __launch_bounds__(256)
stom_prefix_kernel(const uint8_t* __restrict__ dil, int h, int w, uint16_t* __restrict__ pre) {
  const int y = blockIdx.x, f = blockIdx.y;
  const uint8_t* row = dil + (static_cast<size_t>(f) * h + y) * w;
  uint16_t* out = pre + (static_cast<size_t>(f) * h + y) * (w + 1);
  __shared__ int carry, wsum[8];
  if (threadIdx.x == 0) {
    carry = 0;
    out[0] = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int x0 = 0; x0 < w; x0 += blockDim.x) {
    const int x = x0 + threadIdx.x;
    int val = (x < w && row[x]) ? 1 : 0;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, val, o);
      if (lane >= o) val += t;
    }
    if (lane == 31) wsum[warp] = val;
    __syncthreads();
    int before = carry;
    for (int wi = 0; wi < warp; ++wi) before += wsum[wi];
    if (x < w) out[x + 1] = static_cast<uint16_t>(before + val);
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = before + val;
    __syncthreads();
  }
}

// erosion of the dilated mask (= closing of the point mask) and its moments: count, sum x, sum y per frame
__global__ void __launch_bounds__(256)
stom_erode_moments_kernel(const uint16_t* __restrict__ pre, int h, int w, SeRows se, const int32_t* __restrict__ frame_state,
                          int key_idx, unsigned long long* __restrict__ moments) {
  const int y = blockIdx.x, f = blockIdx.y;
  if (f == key_idx || frame_state[f] != 0) return;
  const uint16_t* base = pre + static_cast<size_t>(f) * h * (w + 1);
  unsigned long long cnt = 0, sx = 0;
  for (int x = threadIdx.x; x < w; x += blockDim.x) {
    bool on = true;
    for (int si = 0; si < se.k && on; ++si) {
      const int yy = y + si - se.anchor;
      if (se.j1[si] >= se.j2[si] || yy < 0 || yy >= h) continue;  // outside the image: ignored by cv2's border rule
      const int lo = min(max(x + se.j1[si] - se.anchor, 0), w), hi = min(max(x + se.j2[si] - se.anchor, 0), w);
      const uint16_t* r = base + static_cast<size_t>(yy) * (w + 1);
      on = (r[hi] - r[lo]) == (hi - lo);
    }
    if (on) {
      ++cnt;
      sx += x;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    sx += __shfl_down_sync(0xffffffffu, sx, o);
  }
  if ((threadIdx.x & 31) == 0 && cnt) {
    atomicAdd(&moments[3 * f], cnt);
    atomicAdd(&moments[3 * f + 1], sx);
    atomicAdd(&moments[3 * f + 2], cnt * static_cast<unsigned long long>(y));
  }
}

// first non-transparent pixel of the layer in row-major order (:168-172)
__global__ void __launch_bounds__(256)
stom_first_alpha_kernel(const uint8_t* __restrict__ layer, int npx, unsigned int* __restrict__ first) {
  unsigned int best = 0xffffffffu;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x)
    if (layer[static_cast<size_t>(i) * 4 + 3] != 0) {
      best = i;
      break;  // indices only grow along this thread's stride
    }
  for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_down_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0 && best != 0xffffffffu) atomicMin(first, best);
}

__global__ void stom_points_finalize_kernel(const uint8_t* __restrict__ layer, const unsigned int* __restrict__ first,
                                            const unsigned long long* __restrict__ moments,
                                            const int32_t* __restrict__ frame_state, int t_frames, int key_idx, int h, int w,
                                            b200vit_frame_op* __restrict__ ops) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= t_frames) return;
  if (f == key_idx) {
    write_op(ops + f, B200VIT_FRAME_LAYER, 0, 0, 0, 0, 0, 0, 0, 0u);
    return;
  }
  const unsigned long long cnt = moments[3 * f];
  if (frame_state[f] != 0 || cnt == 0) {  // m00 == 0: a fully transparent layer is composited (:192)
    write_op(ops + f, B200VIT_FRAME_NONE, 0, 0, 0, 0, 0, 0, 0, 0u);
    return;
  }
  uint32_t rgba = 0u;
  const unsigned int fi = *first;
  if (fi != 0xffffffffu) rgba = *reinterpret_cast<const uint32_t*>(layer + static_cast<size_t>(fi) * 4);
  uint32_t a = rgba >> 24;
  a = max(min(a, 148u), 96u);  // :174
  rgba = (rgba & 0x00ffffffu) | (a << 24);
  // cv2.moments: m00 = 255*count, m10 = 255*sum(x), m01 = 255*sum(y) as doubles; int(m10/m00) (:193-195)
  const double m00 = __dmul_rn(255.0, static_cast<double>(cnt));
  const int cx = static_cast<int>(__ddiv_rn(__dmul_rn(255.0, static_cast<double>(moments[3 * f + 1])), m00));
  const int cy = static_cast<int>(__ddiv_rn(__dmul_rn(255.0, static_cast<double>(moments[3 * f + 2])), m00));
  write_op(ops + f, B200VIT_FRAME_CIRCLE, 0, 0, 0, 0, cx, cy, min(h, w) / 20, rgba);
}

// cv2.getStructuringElement(MORPH_ELLIPSE, (k, k)) as one [j1, j2) span per row (see oracle/stom_policy_ref.py)
void ellipse_rows(int k, SeRows& se) {
  std::memset(&se, 0, sizeof(se));
  se.k = k;
  se.anchor = k / 2;
  const int r = k / 2, c = k / 2;
  const double inv_r2 = r ? 1.0 / (static_cast<double>(r) * r) : 0.0;
  for (int i = 0; i < k; ++i) {
    const int dy = i - r;
    if (std::abs(dy) <= r) {
      const int dx = static_cast<int>(std::nearbyint(c * std::sqrt((static_cast<double>(r) * r - static_cast<double>(dy) * dy) * inv_r2)));
      se.j1[i] = static_cast<int16_t>(c - dx > 0 ? c - dx : 0);
      se.j2[i] = static_cast<int16_t>(c + dx + 1 < k ? c + dx + 1 : k);
    }
  }
}

size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

struct PolicyLayout {
  size_t flow_scratch, dil, pre, moments, state, first, total;
};
PolicyLayout policy_layout(int t, int n, int h, int w) {
  PolicyLayout l;
  size_t off = 0;
  l.flow_scratch = off, off += align256(static_cast<size_t>(t) * 6 * n * sizeof(float));
  l.dil = off, off += align256(static_cast<size_t>(t) * h * w);
  l.pre = off, off += align256(static_cast<size_t>(t) * h * (w + 1) * sizeof(uint16_t));
  l.moments = off, off += align256(static_cast<size_t>(t) * 3 * sizeof(unsigned long long));
  l.state = off, off += align256(static_cast<size_t>(t) * sizeof(int32_t));
  l.first = off, off += 256;
  l.total = off;
  return l;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" size_t b200vit_stom_policy_workspace_bytes(int32_t t_frames, int32_t n_points, int32_t h, int32_t w) {
  if (t_frames <= 0 || n_points < 0 || h <= 0 || w <= 0) return 0;
  return policy_layout(t_frames, n_points, h, w).total;
}

extern "C" int b200vit_stom_policy(const float* d_tracks, const uint8_t* d_vis, int32_t t_frames, int32_t n_points,
                                   int32_t key_idx, int32_t mask_shape, int32_t h, int32_t w, const uint8_t* d_layer_rgba,
                                   b200vit_frame_op* d_ops, void* d_workspace, size_t workspace_bytes, b200vit_stream stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = check_arch();
  if (rc) return rc;
  if (t_frames <= 0 || h <= 0 || w <= 0 || w > 65535) return fail(B200VIT_EINVAL, "stom_policy: bad clip shape");
  if (n_points < 0 || n_points > MAX_POINTS) return fail(B200VIT_EINVAL, "stom_policy: n_points must be in [0, 16384]");
  if (key_idx < 0 || key_idx >= t_frames) return fail(B200VIT_EINVAL, "stom_policy: key frame index out of range");
  if (d_ops == nullptr || (n_points > 0 && (d_tracks == nullptr || d_vis == nullptr)))
    return fail(B200VIT_EINVAL, "stom_policy: null pointer");
  if ((reinterpret_cast<uintptr_t>(d_ops) & 3) || (reinterpret_cast<uintptr_t>(d_tracks) & 3))
    return fail(B200VIT_EALIGN, "stom_policy: d_ops / d_tracks must be 4-byte aligned");
  const PolicyLayout lay = policy_layout(t_frames, n_points, h, w);
  if (d_workspace == nullptr || workspace_bytes < lay.total) return fail(B200VIT_EINVAL, "stom_policy: workspace too small");
  if (reinterpret_cast<uintptr_t>(d_workspace) & 255) return fail(B200VIT_EALIGN, "stom_policy: workspace must be 256-byte aligned");
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);

  if (!mask_shape) {
    int p2 = 1;
    while (p2 < n_points) p2 <<= 1;
    const int smem = p2 * static_cast<int>(sizeof(float));
    static DeviceOnce attr;
    if (attr.need()) {
      B200_CUDA_OK(cudaFuncSetAttribute(stom_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        MAX_POINTS * static_cast<int>(sizeof(float))));
      attr.mark();
    }
    stom_flow_kernel<<<t_frames, POLICY_THREADS, smem, stream>>>(d_tracks, d_vis, n_points, key_idx, h, w,
                                                                 reinterpret_cast<float*>(ws + lay.flow_scratch), d_ops);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
  }

  if (d_layer_rgba == nullptr || (reinterpret_cast<uintptr_t>(d_layer_rgba) & 3))
    return fail(B200VIT_EINVAL, "stom_policy: mask shapes need the 4-byte aligned RGBA layer");
  const int k = (h < w ? h : w) / 15;
  if (k < 1 || k > MAX_SE) return fail(B200VIT_EINVAL, "stom_policy: min(h,w)/15 must be in [1,128]");
  if ((h < w ? h : w) / 20 > 127) return fail(B200VIT_EINVAL, "stom_policy: circle radius min(h,w)/20 exceeds 127");
  SeRows se;
  ellipse_rows(k, se);
  uint8_t* dil = ws + lay.dil;
  uint16_t* pre = reinterpret_cast<uint16_t*>(ws + lay.pre);
  unsigned long long* moments = reinterpret_cast<unsigned long long*>(ws + lay.moments);
  int32_t* state = reinterpret_cast<int32_t*>(ws + lay.state);
  unsigned int* first = reinterpret_cast<unsigned int*>(ws + lay.first);
  B200_CUDA_OK(cudaMemsetAsync(dil, 0, static_cast<size_t>(t_frames) * h * w, stream));
  B200_CUDA_OK(cudaMemsetAsync(moments, 0, static_cast<size_t>(t_frames) * 3 * sizeof(unsigned long long), stream));
  B200_CUDA_OK(cudaMemsetAsync(state, 0, static_cast<size_t>(t_frames) * sizeof(int32_t), stream));
  B200_CUDA_OK(cudaMemsetAsync(first, 0xff, sizeof(unsigned int), stream));
  stom_first_alpha_kernel<<<64, 256, 0, stream>>>(d_layer_rgba, h * w, first);
  if (n_points > 0) {
    const dim3 pgrid((n_points + POINTS_PER_BLOCK - 1) / POINTS_PER_BLOCK, t_frames);
    stom_points_kernel<<<pgrid, 256, 0, stream>>>(d_tracks, d_vis, n_points, key_idx, h, w, se, dil, state);
  }
  stom_prefix_kernel<<<dim3(h, t_frames), 256, 0, stream>>>(dil, h, w, pre);
  stom_erode_moments_kernel<<<dim3(h, t_frames), 256, 0, stream>>>(pre, h, w, se, state, key_idx, moments);
  stom_points_finalize_kernel<<<(t_frames + 127) / 128, 128, 0, stream>>>(d_layer_rgba, first, moments, state, t_frames,
                                                                          key_idx, h, w, d_ops);
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}
