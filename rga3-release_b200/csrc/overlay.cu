// STOM visual-prompt overlay + rescale/normalise + 2x14x14 patchify, one pass.
//
//   uint8 frames [T,H,W,3] (+ prompt layer, per-frame ops)  ->  bf16 [M, 3*tp*P*P]
//
// Integer-exact restatement of
//   * Pillow alpha_composite on an opaque frame (the reference composites at
//     /root/reference/model/STOM.py:84-87, :157-160, :204-207),
//   * STOM.warp's forward scatter (:145-155) as a per-destination gather,
//   * STOM.warp_point's cv2.circle stamp (:195-201),
//   * PIL ImageDraw.rectangle outline (visual_prompt_generator.py:102-104),
// followed by HF's fused rescale+normalise (image_processing_backends.py:291-331;
// only 3x256 distinct values -> exact LUT) and the patchify permutation
// (video_processing_qwen2_vl.py:255-272).  HBM-bound: 3 B/px frame + 0/1/4 B/px layer
// in, 6 B/px out; every output element is written once with 16-byte stores.
#include <cuda_bf16.h>

#include <cstring>

#include "internal.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace b200 {
namespace {

constexpr int MAX_FRAMES_PER_LAUNCH = 128;

struct FrameOpDev {
  int32_t mode, sx, sy, cx, cy, r;
  uint8_t zx, zy, pad0, pad1;
  uint32_t rgba;  // little-endian r | g<<8 | b<<16 | a<<24
};

struct OverlayParams {
  int32_t kind;
  const uint8_t* layer;
  int32_t box[4];
  int32_t box_width;
  int32_t circle_r;
  uint32_t palette[256];
  int16_t circle_hw[256];
  const b200vit_frame_op* d_ops;  // device-resident ops of this launch window (b200vit_stom_policy), or nullptr
  FrameOpDev ops[MAX_FRAMES_PER_LAUNCH];
};

struct PatchParams {
  const uint8_t* frames;
  int32_t t, h, w;        // frames in this launch window, image size
  int32_t t_total;        // frames in the clip (for the repeat-last-frame padding)
  int32_t frame_base;     // index of ops[0] / first frame of this launch
  int32_t patch, tps, merge, gh, gw;
  int32_t cols;           // 3*tps*patch*patch
  int64_t row_base;       // first output row of this launch
  int64_t rows;           // rows produced by this launch
};

// bf16 bits of (v - 255*mean_c) / (255*std_c).  Plain global memory, not __constant__: every thread reads a different entry when the
// table is copied to shared memory, and the constant cache serialises divergent addresses (16 % of the kernel in ncu).
__device__ uint16_t g_norm_lut[3 * 256];

__device__ __forceinline__ uint32_t layer_at(const OverlayParams& ov, int h, int w, int y, int x) {
  if (y < 0 || y >= h || x < 0 || x >= w) return 0u;
  if (ov.kind == B200VIT_LAYER_RGBA) {
    return __ldg(reinterpret_cast<const uint32_t*>(ov.layer) + static_cast<size_t>(y) * w + x);
  } else if (ov.kind == B200VIT_LAYER_PALETTE) {
    const uint8_t idx = __ldg(ov.layer + static_cast<size_t>(y) * w + x);
    return idx ? ov.palette[idx] : 0u;
  } else if (ov.kind == B200VIT_LAYER_BOX) {
    // Pillow ImagingDrawRectangle, outline branch: rows t+i / b-i over [l..r]; columns r-i / l+i
    // between rows t+width (included) and b-width+1 (excluded), either direction.
    const int l = ov.box[0], t = ov.box[1], r = ov.box[2], b = ov.box[3], wd = ov.box_width;
    const bool in_x = x >= min(l, r) && x <= max(l, r);
    const bool rows = in_x && ((y >= t && y < t + wd) || (y <= b && y > b - wd));
    const int ya = t + wd, yb = b - wd + 1;
    const bool in_y = (ya < yb) ? (y >= ya && y < yb) : (ya > yb ? (y > yb && y <= ya) : false);
    const bool cols = in_y && ((x <= r && x > r - wd) || (x >= l && x < l + wd));
    return (rows || cols) ? ov.palette[1] : 0u;
  }
  return 0u;
}

// Frame op `rel` of this launch window: from the kernel parameters (host ops, validated by fill_overlay) or from
// device memory (written by b200vit_stom_policy; nothing the host could check, so anything inconsistent
// degrades to "frame untouched").
__device__ __forceinline__ FrameOpDev frame_op_at(const OverlayParams& ov, int rel) {
  if (ov.d_ops == nullptr) return ov.ops[rel];
  const int32_t* s = reinterpret_cast<const int32_t*>(ov.d_ops + rel);
  FrameOpDev d;
  d.mode = __ldg(s), d.sx = __ldg(s + 1), d.sy = __ldg(s + 2);
  d.zx = __ldg(s + 3) ? 1 : 0, d.zy = __ldg(s + 4) ? 1 : 0, d.pad0 = d.pad1 = 0;
  d.cx = __ldg(s + 5), d.cy = __ldg(s + 6), d.r = __ldg(s + 7);
  d.rgba = static_cast<uint32_t>(__ldg(s + 8));  // bytes r,g,b,a little-endian
  const bool ok = (d.mode == B200VIT_FRAME_LAYER && ov.kind != B200VIT_LAYER_NONE) ||
                  (d.mode == B200VIT_FRAME_CIRCLE && d.r == ov.circle_r && d.r >= 0);
  if (!ok) d.mode = B200VIT_FRAME_NONE;
  return d;
}

// RGBA of the prompt layer as seen by destination pixel (y, x) of a frame; alpha 0 = untouched.
__device__ __forceinline__ uint32_t overlay_at(const OverlayParams& ov, const FrameOpDev& op, int h, int w, int y, int x) {
  if (op.mode == B200VIT_FRAME_LAYER) {
    // candidates in row-major SOURCE order; the last one with alpha > 0 wins (STOM.py:149-154)
    const int rb = y - op.sy, cb = x - op.sx;
    const bool er = (y == 0) && op.zy, ec = (x == 0) && op.zx;
    uint32_t v = layer_at(ov, h, w, rb, cb);
    if ((v >> 24) == 0 && ec) v = layer_at(ov, h, w, rb, -1 - op.sx);
    if ((v >> 24) == 0 && er) {
      v = layer_at(ov, h, w, -1 - op.sy, cb);
      if ((v >> 24) == 0 && ec) v = layer_at(ov, h, w, -1 - op.sy, -1 - op.sx);
    }
    return v;
  } else if (op.mode == B200VIT_FRAME_CIRCLE) {
    const int dy = y - op.cy;
    if (dy < -op.r || dy > op.r) return 0u;
    const int hw = ov.circle_hw[dy + op.r];
    const int dx = x - op.cx;
    return (hw >= 0 && dx >= -hw && dx <= hw) ? op.rgba : 0u;
  }
  return 0u;
}

// Pillow AlphaComposite.c with dst alpha = 255 (see oracle/overlay_ref.py)
__device__ __forceinline__ uint32_t composite_ch(uint32_t d, uint32_t s, uint32_t a) {
  const uint32_t t = s * (a * 128u) + d * ((255u - a) * 128u) + (0x80u << 7);
  return (((t >> 8) + t) >> 8) >> 7;
}

template <bool HAS_OVERLAY>
__global__ void __launch_bounds__(256)
overlay_patchify_kernel(const __grid_constant__ PatchParams p, const __grid_constant__ OverlayParams ov,
                        __nv_bfloat16* __restrict__ out) {
  griddep_launch_dependents();
  __shared__ uint16_t lut[3 * 256];
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) lut[i] = g_norm_lut[i];
  __syncthreads();
  griddep_wait();  // the output buffer may still be read by the previous forward's patch-embed GEMM
  const int chunks = p.cols >> 3;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= p.rows * chunks) return;
  const int64_t row = gid / chunks;
  const int col0 = static_cast<int>(gid % chunks) * 8;
  // row -> (t, bh, bw, mh, mw)
  const int mm = p.merge * p.merge;
  const int gw2 = p.gw / p.merge, gh2 = p.gh / p.merge;
  const int64_t grow = row + p.row_base;
  const int mi = static_cast<int>(grow % mm);
  int64_t rest = grow / mm;
  const int bw = static_cast<int>(rest % gw2);
  rest /= gw2;
  const int bh = static_cast<int>(rest % gh2);
  const int tt = static_cast<int>(rest / gh2);
  const int y0 = (bh * p.merge + mi / p.merge) * p.patch;
  const int x0 = (bw * p.merge + mi % p.merge) * p.patch;
  const int pp = p.patch * p.patch;
  uint32_t packed[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = col0 + j;
    const int c = col / (p.tps * pp);
    const int r2 = col % (p.tps * pp);
    const int tp = r2 / pp;
    const int ph = (r2 % pp) / p.patch;
    const int pw = r2 % p.patch;
    int f = tt * p.tps + tp;
    f = min(f, p.t_total - 1);  // odd T: repeat the last frame (HF videoproc :245-249)
    const int y = y0 + ph, x = x0 + pw;
    uint32_t d = __ldg(p.frames + ((static_cast<size_t>(f - p.frame_base) * p.h + y) * p.w + x) * 3 + c);
    if (HAS_OVERLAY) {
      const uint32_t s = overlay_at(ov, frame_op_at(ov, f - p.frame_base), p.h, p.w, y, x);
      const uint32_t a = s >> 24;
      if (a) d = composite_ch(d, (s >> (8 * c)) & 0xffu, a);
    }
    const uint32_t bits = lut[c * 256 + d];
    if (j & 1)
      packed[j >> 1] |= bits << 16;
    else
      packed[j >> 1] = bits;
  }
  *reinterpret_cast<uint4*>(out + row * p.cols + col0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
}

// Strip version for the Qwen2.5-VL geometry (patch 14, temporal 2, merge 2): a block owns G consecutive
// merge groups of one (t, merge-row) strip -- 28 rows x 28G pixels x 2 frames in, 4G consecutive patch
// rows out.  Pixels are visited in raster order (byte loads of consecutive lanes fall in the same
// sectors, the overlay is evaluated once per pixel instead of once per channel), values are scattered
// into the patch layout in shared memory and leave as one contiguous, 16-byte-vectorised block.
constexpr int SP = 14, STPS = 2, SMG = 2, SG = 2;           // geometry, groups per block
constexpr int SCOLS = 3 * STPS * SP * SP;                   // 1176
constexpr int STRIP_OUT_BYTES = SG * SMG * SMG * SCOLS * 2; // 18,816 B  patch-layout bf16 block
constexpr int STRIP_RH = SP * SMG;                          // 28 pixel rows per strip
constexpr int STRIP_ROW_BYTES = SG * STRIP_RH * 3;          // 168 B of RGB per staged pixel row
constexpr int STRIP_IN_BYTES = STPS * STRIP_RH * STRIP_ROW_BYTES;  // 9,408 B  staged uint8 pixels
constexpr int STRIP_SMEM = STRIP_OUT_BYTES + STRIP_IN_BYTES;

template <bool HAS_OVERLAY>
__global__ void __launch_bounds__(256)
overlay_patchify_strip_kernel(const __grid_constant__ PatchParams p, const __grid_constant__ OverlayParams ov,
                              __nv_bfloat16* __restrict__ out) {
  griddep_launch_dependents();
  extern __shared__ __align__(16) uint8_t strip_raw[];
  uint16_t* out_s = reinterpret_cast<uint16_t*>(strip_raw);
  uint8_t* in_s = strip_raw + STRIP_OUT_BYTES;
  __shared__ uint16_t lut[3 * 256];
  constexpr int RH = STRIP_RH;
  const int gw2 = p.gw / SMG, gh2 = p.gh / SMG;
  const int bw0 = blockIdx.x * SG;
  const int groups = min(SG, gw2 - bw0);
  const int bh = blockIdx.y;
  const int tt = blockIdx.z + static_cast<int>(p.row_base / (static_cast<int64_t>(p.gh) * p.gw));  // absolute temporal index
  const int rw = groups * RH;           // pixel columns of this block
  // stage the block's pixels with 4-byte coalesced loads (rows of rw*3 contiguous bytes; every row start is 4-byte
  // aligned because W and the block's first column are multiples of 28): all loads of a thread are independent
  {
    // every thread owns one word column of 14 staged rows: issue all 14 loads, then store (the kernel is latency-bound:
    // 10 MB in, the 19 MB result stays in L2; a load->store loop left one request in flight per warp)
    const int row_words = rw * 3 / 4;
    const int y0 = bh * RH, x0 = bw0 * RH;
    const int wi = threadIdx.x & 63;
    constexpr int NR = STPS * RH / 4;  // 14 rows per thread
    uint32_t tmp[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const int r = (threadIdx.x >> 6) + 4 * k;
      const int tp = r / RH, yy = r - tp * RH;
      const int f = min(tt * STPS + tp, p.t_total - 1);  // odd T: repeat the last frame (HF videoproc :245-249)
      const uint32_t* src = reinterpret_cast<const uint32_t*>(
          p.frames + ((static_cast<size_t>(f - p.frame_base) * p.h + (y0 + yy)) * p.w + x0) * 3);
      tmp[k] = wi < row_words ? __ldg(src + wi) : 0u;
    }
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const int r = (threadIdx.x >> 6) + 4 * k;
      if (wi < row_words) reinterpret_cast<uint32_t*>(in_s + r * STRIP_ROW_BYTES)[wi] = tmp[k];
    }
  }
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) lut[i] = g_norm_lut[i];
  FrameOpDev fop[STPS];  // the ops of this block's STPS frames
  if (HAS_OVERLAY) {
#pragma unroll
    for (int tp = 0; tp < STPS; ++tp) fop[tp] = frame_op_at(ov, min(tt * STPS + tp, p.t_total - 1) - p.frame_base);
  }
  __syncthreads();
  griddep_wait();  // the output buffer may still be read by the previous forward's patch-embed GEMM
  // one thread = one pixel column (64 lanes, rw <= 56 used) x every 4th pixel row: no index divisions in the loop
  const int xx = threadIdx.x & 63;
  if (xx < rw) {
    const int x = bw0 * RH + xx;
    const int g = xx / RH, xg = xx - g * RH;
    const int col_part = (g * SMG * SMG + xg / SP) * SCOLS + (xg % SP);
    constexpr int NP = RH / 4;  // 7 pixel rows per thread and frame
    uint32_t sv[STPS][NP];
    if (HAS_OVERLAY) {  // all layer look-ups of this thread first: independent loads in flight together
#pragma unroll
      for (int tp = 0; tp < STPS; ++tp)
#pragma unroll
        for (int k = 0; k < NP; ++k)
          sv[tp][k] = overlay_at(ov, fop[tp], p.h, p.w, bh * RH + (threadIdx.x >> 6) + 4 * k, x);
    }
#pragma unroll
    for (int tp = 0; tp < STPS; ++tp) {
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const int yy = (threadIdx.x >> 6) + 4 * k;
        const uint8_t* px = in_s + (tp * RH + yy) * STRIP_ROW_BYTES + xx * 3;
        uint32_t d0 = px[0], d1 = px[1], d2 = px[2];
        if (HAS_OVERLAY) {
          const uint32_t a = sv[tp][k] >> 24;
          if (a) {
            d0 = composite_ch(d0, sv[tp][k] & 0xffu, a);
            d1 = composite_ch(d1, (sv[tp][k] >> 8) & 0xffu, a);
            d2 = composite_ch(d2, (sv[tp][k] >> 16) & 0xffu, a);
          }
        }
        const int half = yy >= SP ? 1 : 0;  // mi = (yy / SP) * SMG + xg / SP
        const int ph = yy - half * SP;
        uint16_t* dst = out_s + col_part + half * SMG * SCOLS + tp * SP * SP + ph * SP;
        dst[0] = lut[d0];
        dst[STPS * SP * SP] = lut[256 + d1];
        dst[2 * STPS * SP * SP] = lut[512 + d2];
      }
    }
  }
  __syncthreads();
  const int64_t row0 = (static_cast<int64_t>(blockIdx.z) * gh2 + bh) * gw2 * SMG * SMG + static_cast<int64_t>(bw0) * SMG * SMG;
  uint4* gdst = reinterpret_cast<uint4*>(out + row0 * SCOLS);
  const uint4* ssrc = reinterpret_cast<const uint4*>(out_s);
  const int nvec = groups * SMG * SMG * SCOLS * 2 / 16;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) gdst[i] = ssrc[i];
}

__global__ void __launch_bounds__(256)
overlay_composite_kernel(const __grid_constant__ PatchParams p, const __grid_constant__ OverlayParams ov,
                         uint8_t* __restrict__ out) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t npx = static_cast<int64_t>(p.t) * p.h * p.w;
  if (gid >= npx) return;
  const int x = static_cast<int>(gid % p.w);
  const int y = static_cast<int>((gid / p.w) % p.h);
  const int f = static_cast<int>(gid / (static_cast<int64_t>(p.w) * p.h));
  const uint32_t s = overlay_at(ov, frame_op_at(ov, f), p.h, p.w, y, x);
  const uint32_t a = s >> 24;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint32_t d = p.frames[gid * 3 + c];
    if (a) d = composite_ch(d, (s >> (8 * c)) & 0xffu, a);
    out[gid * 3 + c] = static_cast<uint8_t>(d);
  }
}

uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  const uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return static_cast<uint16_t>(u >> 16);
}

int upload_lut() {
  static DeviceOnce done;  // g_norm_lut is a __device__ symbol: one copy per device
  if (!done.need()) return 0;
  // HF: mean_t = float32(mean) * 255, std_t = float32(std) * 255; (float(v) - mean_t) / std_t in fp32
  const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
  const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
  uint16_t lut[3 * 256];
  for (int c = 0; c < 3; ++c) {
    volatile float m = mean[c] * 255.0f;
    volatile float s = stdv[c] * 255.0f;
    for (int v = 0; v < 256; ++v) {
      volatile float d = static_cast<float>(v) - m;
      volatile float q = d / s;
      lut[c * 256 + v] = f32_to_bf16_rne(q);
    }
  }
  B200_CUDA_OK(cudaMemcpyToSymbol(g_norm_lut, lut, sizeof(lut)));
  done.mark();
  return 0;
}

// OpenCV drawing.cpp Circle(): integer midpoint circle, filled -> per-row half widths.
void circle_halfwidths(int r, int16_t* hw) {
  for (int i = 0; i < 2 * r + 1; ++i) hw[i] = -1;
  int err = 0, dx = r, dy = 0, plus = 1, minus = (r << 1) - 1;
  while (dx >= dy) {
    const int rows[2] = {dy, dx}, half[2] = {dx, dy};
    for (int k = 0; k < 2; ++k)
      for (int sgn = -1; sgn <= 1; sgn += 2) {
        const int i = r + sgn * rows[k];
        if (half[k] > hw[i]) hw[i] = static_cast<int16_t>(half[k]);
      }
    dy++;
    err += plus;
    plus += 2;
    const int mask = (err <= 0) - 1;
    err -= minus & mask;
    dx += mask;
    minus -= mask & 2;
  }
}

int fill_overlay(OverlayParams& o, const b200vit_overlay* ov, int f0, int nf, int t_total) {
  std::memset(&o, 0, sizeof(o));
  if (ov == nullptr) return 0;
  o.kind = ov->kind;
  o.layer = ov->d_layer;
  if ((ov->kind == B200VIT_LAYER_RGBA || ov->kind == B200VIT_LAYER_PALETTE) && ov->d_layer == nullptr)
    return fail(B200VIT_EINVAL, "overlay: layer kind needs d_layer");
  if (ov->kind == B200VIT_LAYER_RGBA && (reinterpret_cast<uintptr_t>(ov->d_layer) & 3))
    return fail(B200VIT_EALIGN, "overlay: RGBA layer must be 4-byte aligned");
  for (int i = 0; i < 4; ++i) o.box[i] = ov->box[i];
  o.box_width = ov->box_width < 1 ? 1 : ov->box_width;
  for (int i = 0; i < 256; ++i)
    o.palette[i] = ov->palette[i][0] | (ov->palette[i][1] << 8) | (ov->palette[i][2] << 16) |
                   (static_cast<uint32_t>(ov->palette[i][3]) << 24);
  o.circle_r = -1;
  if (ov->h_ops == nullptr && ov->d_ops != nullptr) {
    // device-resident ops: the kernels read (and sanity-check) them; only the shared circle radius is host knowledge
    if (ov->d_ops_circle_r > 127) return fail(B200VIT_EINVAL, "overlay: circle radius must be in [0,127]");
    if (reinterpret_cast<uintptr_t>(ov->d_ops) & 3) return fail(B200VIT_EALIGN, "overlay: d_ops must be 4-byte aligned");
    o.d_ops = ov->d_ops + f0;
    o.circle_r = ov->d_ops_circle_r >= 0 ? ov->d_ops_circle_r : -1;
    if (o.circle_r >= 0) circle_halfwidths(o.circle_r, o.circle_hw);
    return 0;
  }
  for (int i = 0; i < nf; ++i) {
    FrameOpDev& d = o.ops[i];
    if (ov->h_ops == nullptr || f0 + i >= t_total) {
      d.mode = B200VIT_FRAME_NONE;
      continue;
    }
    const b200vit_frame_op& s = ov->h_ops[f0 + i];
    d.mode = s.mode, d.sx = s.sx, d.sy = s.sy, d.cx = s.cx, d.cy = s.cy, d.r = s.r;
    d.zx = s.zx ? 1 : 0, d.zy = s.zy ? 1 : 0;
    d.rgba = s.rgba[0] | (s.rgba[1] << 8) | (s.rgba[2] << 16) | (static_cast<uint32_t>(s.rgba[3]) << 24);
    if (s.mode == B200VIT_FRAME_CIRCLE) {
      if (s.r < 0 || s.r > 127) return fail(B200VIT_EINVAL, "overlay: circle radius must be in [0,127]");
      if (o.circle_r >= 0 && o.circle_r != s.r)
        return fail(B200VIT_EINVAL, "overlay: all circle stamps of a clip share one radius (min(h,w)//20)");
      o.circle_r = s.r;
    } else if (s.mode == B200VIT_FRAME_LAYER && ov->kind == B200VIT_LAYER_NONE) {
      return fail(B200VIT_EINVAL, "overlay: frame op uses the layer but kind is NONE");
    } else if (s.mode < 0 || s.mode > 2) {
      return fail(B200VIT_EINVAL, "overlay: unknown frame mode");
    }
  }
  if (o.circle_r >= 0) circle_halfwidths(o.circle_r, o.circle_hw);
  return 0;
}

}  // namespace

// out_bf16 != nullptr: patchified bf16 matrix; out_u8 != nullptr: composited frames.
int launch_overlay_patchify(const b200vit_frames& fr, const b200vit_overlay* ov, int patch, int tps, int merge,
                            void* out_bf16, uint8_t* out_u8, cudaStream_t stream) {
  if (fr.d_frames == nullptr || fr.t <= 0 || fr.h <= 0 || fr.w <= 0) return fail(B200VIT_EINVAL, "frames: empty clip");
  const int unit = patch * merge;
  if (out_bf16 && (fr.h % unit || fr.w % unit))
    return fail(B200VIT_EINVAL, "frames: H and W must be multiples of patch*merge (resize is out of scope)");
  const int cols = 3 * tps * patch * patch;
  if (out_bf16 && cols % 8) return fail(B200VIT_EINVAL, "patchify: 3*tps*patch^2 must be a multiple of 8");
  if (out_bf16 && (reinterpret_cast<uintptr_t>(out_bf16) & 15)) return fail(B200VIT_EALIGN, "patchify: out must be 16-byte aligned");
  int rc = upload_lut();
  if (rc) return rc;
  const int gh = fr.h / patch, gw = fr.w / patch;
  const int t_pad = (fr.t + tps - 1) / tps * tps;
  // frames are processed in windows of whole temporal patches so per-frame ops fit the kernel parameters
  const int win = MAX_FRAMES_PER_LAUNCH / tps * tps;
  for (int f0 = 0; f0 < t_pad; f0 += win) {
    const int nf_pad = (t_pad - f0 < win) ? (t_pad - f0) : win;      // padded frames covered
    const int nf = (fr.t - f0 < nf_pad) ? (fr.t - f0) : nf_pad;      // real frames available
    OverlayParams o;
    rc = fill_overlay(o, ov, f0, nf, fr.t);
    if (rc) return rc;
    PatchParams p;
    p.frames = fr.d_frames + static_cast<size_t>(f0) * fr.h * fr.w * 3;
    p.t = nf, p.h = fr.h, p.w = fr.w, p.t_total = fr.t, p.frame_base = f0;
    p.patch = patch, p.tps = tps, p.merge = merge, p.gh = gh, p.gw = gw, p.cols = cols;
    p.row_base = static_cast<int64_t>(f0 / tps) * gh * gw;
    p.rows = static_cast<int64_t>(nf_pad / tps) * gh * gw;
    if (out_bf16) {
      const int64_t threads = p.rows * (cols / 8);
      const int grid = static_cast<int>((threads + 255) / 256);
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(out_bf16) + p.row_base * cols;
      const bool has_ov = ov != nullptr && (ov->h_ops != nullptr || ov->d_ops != nullptr);
      if (patch == SP && tps == STPS && merge == SMG && (reinterpret_cast<uintptr_t>(p.frames) & 3) == 0) {
        static DeviceOnce attr_set;
        if (attr_set.need()) {
          B200_CUDA_OK(cudaFuncSetAttribute(overlay_patchify_strip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, STRIP_SMEM));
          B200_CUDA_OK(cudaFuncSetAttribute(overlay_patchify_strip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, STRIP_SMEM));
          attr_set.mark();
        }
        const dim3 sgrid((gw / SMG + SG - 1) / SG, gh / SMG, nf_pad / tps);
        if (has_ov)
          B200_CUDA_OK(launch_kernel(overlay_patchify_strip_kernel<true>, sgrid, dim3(256), STRIP_SMEM, stream, 1, p, o, dst));
        else
          B200_CUDA_OK(launch_kernel(overlay_patchify_strip_kernel<false>, sgrid, dim3(256), STRIP_SMEM, stream, 1, p, o, dst));
      } else if (has_ov) {
        B200_CUDA_OK(launch_kernel(overlay_patchify_kernel<true>, dim3(grid), dim3(256), 0, stream, 1, p, o, dst));
      } else {
        B200_CUDA_OK(launch_kernel(overlay_patchify_kernel<false>, dim3(grid), dim3(256), 0, stream, 1, p, o, dst));
      }
    }
    if (out_u8) {
      const int64_t npx = static_cast<int64_t>(nf) * fr.h * fr.w;
      const int grid = static_cast<int>((npx + 255) / 256);
      p.frame_base = 0;
      overlay_composite_kernel<<<grid, 256, 0, stream>>>(p, o, out_u8 + static_cast<size_t>(f0) * fr.h * fr.w * 3);
    }
    B200_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

}  // namespace b200
