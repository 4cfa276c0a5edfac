// Prompt-layer rasterisers: the mask and scribble visual prompts of
// /root/reference/utils/visual_prompt_generator.py (draw_mask :268-274 = ImageDraw.polygon(fill) per contour,
// draw_scribble :230-252 = 1000 * scale straight ImageDraw.line(width) segments along a cubic Bezier), drawn into a
// 1-byte palette layer on the GPU with exactly the pixel coverage Pillow produces.
//
// The arithmetic that decides coverage lives in Pillow's C library (src/libImaging/Draw.c, pinned 11.1.0 by the
// reference, 12.2.0 in this image; not part of /root/reference).  It is restated here from Pillow's published
// algorithm and the behaviour of the installed build (fuzzed against live PIL in tests/test_raster_cpu.py and
// tests/test_gpu_raster.py):
//   * ImagingDrawPolygon (fill): vertices truncated to int, one Edge per side (consecutive horizontal sides that keep
//     their direction are merged), closing edge added when the ring is open.
//   * polygon_generic: for every scanline the x of every edge that spans it -- float32 (y - y0) * dx + x0, two
//     roundings -- is collected; an edge that ENDS on the scanline (not on the polygon's last one) contributes its x
//     twice; an edge with an endpoint on the scanline may have its x pulled towards a neighbouring edge that shares
//     the corner ("connect discontiguous corners"); horizontal edges are drawn directly.  The sorted list is consumed
//     in pairs: pixels ROUND_UP(x[2m]) .. ROUND_DOWN(x[2m+1]).
//   * ImagingDrawWideLine: endpoints truncated to int, a 4-vertex polygon from double-precision offsets rounded with
//     ROUND_UP / ROUND_DOWN, filled by polygon_generic; width <= 1 is Bresenham (line32) plus the end point.
// Every pixel written gets the same palette index, so polygons, segments and scanlines are independent: one CTA per
// (polygon, scanline).  The pairing of the SORTED crossings is evaluated without sorting: with a(x) = #{crossings with
// ROUND_UP <= x} and b(x) = #{crossings with ROUND_DOWN < x} (both prefixes of the sorted order, because the two
// roundings are monotone), pixel x is inside some pair iff an even p exists with max(b-1, 0) <= p <= min(a-1, n-2).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "internal.h"

namespace b200 {
namespace {

struct REdge {  // Pillow's Edge (d is unused by the fill)
  int32_t x0, y0, xmin, ymin, xmax, ymax;
  float dx;
};
struct RPoly {
  int32_t edge_begin, edge_count;
  int32_t ymin, ymax;      // scanline range of polygon_generic, clamped to [0, h]
  int32_t bx0, bx1;        // pixel columns any span can touch, clamped to [0, w-1]
  int32_t task_begin;      // first (polygon, row) task of this polygon
  int32_t row0;            // first row that has a task
};

constexpr int kMaxCrossings = 2048;  // per (polygon, scanline); checked on the host
constexpr int kMaxHlines = 2048;
constexpr int kCoordLimit = 1 << 22;

// ---- Pillow's rounding macros.  `(f) + 0.5F` is float arithmetic for a float argument, double for a double one;
// the negative branch goes through fabs() and is always double.
inline int round_up_d(double f) { return static_cast<int>(f >= 0.0 ? std::floor(f + 0.5) : -std::floor(std::fabs(f) + 0.5)); }
inline int round_down_d(double f) { return static_cast<int>(f >= 0.0 ? std::ceil(f - 0.5) : -std::ceil(std::fabs(f) - 0.5)); }

__device__ __forceinline__ int round_up_f(float f) {
  return f >= 0.0f ? static_cast<int>(floorf(__fadd_rn(f, 0.5f))) : -static_cast<int>(floor(static_cast<double>(fabsf(f)) + 0.5));
}
__device__ __forceinline__ int round_down_f(float f) {
  return f >= 0.0f ? static_cast<int>(ceilf(__fsub_rn(f, 0.5f))) : -static_cast<int>(ceil(static_cast<double>(fabsf(f)) - 0.5));
}
// x of edge e on scanline y: cvtsi2ss, mulss, cvtsi2ss, addss -- no contraction
__device__ __forceinline__ float edge_x(const REdge& e, int y) {
  return __fadd_rn(__fmul_rn(__int2float_rn(y - e.y0), e.dx), __int2float_rn(e.x0));
}

__global__ void __launch_bounds__(128)
raster_kernel(const REdge* __restrict__ edges, const RPoly* __restrict__ polys, int n_polys, int n_tasks, int w,
              uint8_t index, uint8_t* __restrict__ layer) {
  __shared__ int s_up[kMaxCrossings];    // ROUND_UP of every crossing
  __shared__ int s_down[kMaxCrossings];  // ROUND_DOWN of every crossing
  __shared__ int s_h0[kMaxHlines], s_h1[kMaxHlines];
  __shared__ int s_n, s_nh;
  const int task = blockIdx.x;
  if (task >= n_tasks) return;
  // polygon of this task: last polygon whose task_begin <= task
  int lo = 0, hi = n_polys - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (polys[mid].task_begin <= task) lo = mid; else hi = mid - 1;
  }
  const RPoly P = polys[lo];
  const int y = P.row0 + (task - P.task_begin);
  const REdge* E = edges + P.edge_begin;
  if (threadIdx.x == 0) s_n = 0, s_nh = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < P.edge_count; i += blockDim.x) {
    const REdge cur = E[i];
    if (cur.ymin == cur.ymax) {  // horizontal edge: hline(xmin .. xmax) on its own row
      if (cur.ymin == y) {
        const int slot = atomicAdd(&s_nh, 1);
        if (slot < kMaxHlines) s_h0[slot] = cur.xmin, s_h1[slot] = cur.xmax;
      }
      continue;
    }
    if (y < cur.ymin || y > cur.ymax) continue;
    float x = edge_x(cur, y);
    int copies = 1;
    if (y == cur.ymax && y < P.ymax) {
      copies = 2;  // "needed to draw consistent polygons"
    } else if ((y == cur.ymin || y == cur.ymax) && cur.dx != 0.0f) {
      // connect discontiguous corners: the first earlier edge that also has an endpoint on this scanline, is not
      // vertical, meets this edge at the same rounded x and exists on the adjacent scanline decides
      const int adj = (y == cur.ymax) ? y - 1 : y + 1;
      const float rx = roundf(x);
      for (int k = 0; k < i; ++k) {
        const REdge o = E[k];
        if (y != o.ymin && y != o.ymax) continue;
        if (o.ymin == o.ymax || o.dx == 0.0f) continue;
        if (rx != roundf(edge_x(o, y))) continue;
        if (adj < o.ymin || adj > o.ymax) continue;
        const float ax = edge_x(cur, adj), ao = edge_x(o, adj);
        if (x > __fadd_rn(ax, 1.0f) && x > __fadd_rn(ao, 1.0f))
          x = __fadd_rn(roundf(fmaxf(ax, ao)), 1.0f);
        else if (__fsub_rn(ax, 1.0f) > x && __fsub_rn(ao, 1.0f) > x)
          x = __fsub_rn(roundf(fminf(ax, ao)), 1.0f);
        break;
      }
    }
    const int slot = atomicAdd(&s_n, copies);
    const int up = round_up_f(x), down = round_down_f(x);
    for (int c = 0; c < copies; ++c)
      if (slot + c < kMaxCrossings) s_up[slot + c] = up, s_down[slot + c] = down;
  }
  __syncthreads();
  const int n = min(s_n, kMaxCrossings), nh = min(s_nh, kMaxHlines);
  if (y < 0) return;
  uint8_t* row = layer + static_cast<size_t>(y) * w;
  for (int x = P.bx0 + threadIdx.x; x <= P.bx1; x += blockDim.x) {
    int a = 0, b = 0;
    for (int i = 0; i < n; ++i) {
      a += s_up[i] <= x;
      b += s_down[i] < x;
    }
    const int plo = max(b - 1, 0), phi = min(a - 1, n - 2);
    bool covered = phi >= plo && ((plo & 1) == 0 || plo + 1 <= phi);
    for (int i = 0; i < nh && !covered; ++i) covered = s_h0[i] <= x && x <= s_h1[i];
    if (covered) row[x] = index;
  }
}

// ---- host: Pillow's edge construction
void add_edge(std::vector<REdge>& out, int x0, int y0, int x1, int y1) {
  REdge e;
  e.xmin = std::min(x0, x1), e.xmax = std::max(x0, x1);
  e.ymin = std::min(y0, y1), e.ymax = std::max(y0, y1);
  e.dx = (y0 == y1) ? 0.0f : static_cast<float>(x1 - x0) / static_cast<float>(y1 - y0);
  e.x0 = x0, e.y0 = y0;
  out.push_back(e);
}

struct Builder {
  int h, w;
  std::vector<REdge> edges;
  std::vector<RPoly> polys;
  int n_tasks = 0;
  std::string err;

  bool coord_ok(int v) const { return v > -kCoordLimit && v < kCoordLimit; }

  // closes the polygon whose edges are edges[begin ..]: scanline range, column range, task range, capacity checks
  bool finish(int begin) {
    const int count = static_cast<int>(edges.size()) - begin;
    if (count <= 0) return true;
    int ymin = h - 1, ymax = 0, xmin = w - 1, xmax = 0;
    for (int i = begin; i < begin + count; ++i) {
      const REdge& e = edges[i];
      ymin = std::min(ymin, e.ymin), ymax = std::max(ymax, e.ymax);
      xmin = std::min(xmin, e.xmin), xmax = std::max(xmax, e.xmax);
    }
    RPoly p;
    p.edge_begin = begin, p.edge_count = count;
    p.ymin = std::max(ymin, 0), p.ymax = std::min(ymax, h);
    // every crossing lies in [xmin, xmax]; a corner adjustment moves one to (a neighbouring edge's x) +- 1
    p.bx0 = std::max(xmin - 2, 0), p.bx1 = std::min(xmax + 2, w - 1);
    p.row0 = p.ymin;
    const int rows = std::min(p.ymax, h - 1) - p.ymin + 1;
    if (rows <= 0) {
      edges.resize(begin);
      return true;
    }
    // capacity: crossings (two per edge at most) and horizontal edges per scanline
    std::vector<int> cross(rows + 1, 0), hl(rows + 1, 0);
    for (int i = begin; i < begin + count; ++i) {
      const REdge& e = edges[i];
      const int a = std::max(e.ymin, p.ymin) - p.ymin, b = std::min(e.ymax, p.ymin + rows - 1) - p.ymin;
      if (a > b) continue;
      if (e.ymin == e.ymax) hl[a] += 1;
      else for (int r = a; r <= b; ++r) cross[r] += 2;
    }
    for (int r = 0; r < rows; ++r)
      if (cross[r] > kMaxCrossings || hl[r] > kMaxHlines) {
        err = "raster: more than 2048 edge crossings (or horizontal edges) on one scanline of one polygon";
        return false;
      }
    p.task_begin = n_tasks;
    n_tasks += rows;
    polys.push_back(p);
    return true;
  }

  // ImagingDrawPolygon, fill branch; xy = count (x, y) pairs already truncated to int
  bool polygon(const int* xy, int count) {
    if (count <= 0) return true;
    for (int i = 0; i < 2 * count; ++i)
      if (!coord_ok(xy[i])) {
        err = "raster: coordinate out of range";
        return false;
      }
    const int begin = static_cast<int>(edges.size());
    int i = 0;
    for (i = 0; i < count - 1; ++i) {
      const int x0 = xy[i * 2], y0 = xy[i * 2 + 1], x1 = xy[i * 2 + 2], y1 = xy[i * 2 + 3];
      if (y0 == y1 && i != 0 && y0 == xy[i * 2 - 1] && static_cast<int>(edges.size()) > begin) {
        // a horizontal side right after another horizontal side: extend it when both run the same way
        REdge& last = edges.back();
        if (x1 > x0 && x0 > xy[i * 2 - 2]) {
          last.xmax = x1;
          continue;
        }
        if (x1 < x0 && x0 < xy[i * 2 - 2]) {
          last.xmin = x1;
          continue;
        }
      }
      add_edge(edges, x0, y0, x1, y1);
    }
    if (xy[i * 2] != xy[0] || xy[i * 2 + 1] != xy[1]) add_edge(edges, xy[i * 2], xy[i * 2 + 1], xy[0], xy[1]);
    return finish(begin);
  }

  void point(int x, int y) {  // draw->point as a degenerate horizontal edge (hline clips it)
    const int begin = static_cast<int>(edges.size());
    add_edge(edges, x, y, x, y);
    finish(begin);
  }

  // line32: Bresenham without its last pixel.  Pixels that follow each other on one row become one horizontal edge.
  void thin_line(int x0, int y0, int x1, int y1) {
    const int begin = static_cast<int>(edges.size());
    int dx = x1 - x0, dy = y1 - y0, xs = 1, ys = 1;
    if (dx < 0) dx = -dx, xs = -1;
    if (dy < 0) dy = -dy, ys = -1;
    bool open = false;
    int ry = 0, ra = 0, rb = 0;
    auto flush = [&]() {
      if (open) add_edge(edges, ra, ry, rb, ry);
      open = false;
    };
    auto px = [&](int x, int y) {
      if (open && y == ry && (x == rb + 1 || x == ra - 1)) {
        ra = std::min(ra, x), rb = std::max(rb, x);
        return;
      }
      flush();
      open = true, ry = y, ra = rb = x;
    };
    if (dx == 0) {
      for (int i = 0; i < dy; ++i, y0 += ys) px(x0, y0);
    } else if (dy == 0) {
      for (int i = 0; i < dx; ++i, x0 += xs) px(x0, y0);
    } else if (dx > dy) {
      const int n = dx;
      dy += dy;
      int e = dy - dx;
      dx += dx;
      for (int i = 0; i < n; ++i) {
        px(x0, y0);
        if (e >= 0) y0 += ys, e -= dx;
        e += dy;
        x0 += xs;
      }
    } else {
      const int n = dy;
      dx += dx;
      int e = dx - dy;
      dy += dy;
      for (int i = 0; i < n; ++i) {
        px(x0, y0);
        if (e >= 0) x0 += xs, e -= dy;
        e += dx;
        y0 += ys;
      }
    }
    flush();
    finish(begin);
  }

  // _draw_lines for one two-point call: draw.line([(x0, y0), (x1, y1)], width)
  bool line(double fx0, double fy0, double fx1, double fy1, int width) {
    const double lim = static_cast<double>(kCoordLimit);
    if (!(std::fabs(fx0) < lim && std::fabs(fy0) < lim && std::fabs(fx1) < lim && std::fabs(fy1) < lim)) {
      err = "raster: coordinate out of range";
      return false;
    }
    const int x0 = static_cast<int>(fx0), y0 = static_cast<int>(fy0), x1 = static_cast<int>(fx1), y1 = static_cast<int>(fy1);
    if (width <= 1) {
      thin_line(x0, y0, x1, y1);
      point(x1, y1);  // "draw last point"
      return true;
    }
    const int dx = x1 - x0, dy = y1 - y0;
    if (dx == 0 && dy == 0) {
      point(x0, y0);
      return true;
    }
    // ImagingDrawWideLine
    const double big = std::hypot(static_cast<double>(dx), static_cast<double>(dy));
    const double small = (width - 1) / 2.0;
    const double ratio_max = round_up_d(small) / big, ratio_min = round_down_d(small) / big;
    const int dxmin = round_down_d(ratio_min * dy), dxmax = round_down_d(ratio_max * dy);
    const int dymin = round_up_d(ratio_min * dx), dymax = round_up_d(ratio_max * dx);
    const int v[4][2] = {{x0 - dxmin, y0 + dymax}, {x1 - dxmin, y1 + dymax}, {x1 + dxmax, y1 - dymin}, {x0 + dxmax, y0 - dymin}};
    const int begin = static_cast<int>(edges.size());
    for (int i = 0; i < 4; ++i) add_edge(edges, v[i][0], v[i][1], v[(i + 1) & 3][0], v[(i + 1) & 3][1]);
    return finish(begin);
  }
};

int launch(Builder& b, uint8_t index, uint8_t* d_layer, cudaStream_t stream) {
  if (b.polys.empty()) return 0;
  REdge* d_edges = nullptr;
  RPoly* d_polys = nullptr;
  B200_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&d_edges), b.edges.size() * sizeof(REdge), stream));
  B200_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&d_polys), b.polys.size() * sizeof(RPoly), stream));
  // pageable sources: the copies complete (for the host) before these calls return, so the vectors may die
  B200_CUDA_OK(cudaMemcpyAsync(d_edges, b.edges.data(), b.edges.size() * sizeof(REdge), cudaMemcpyHostToDevice, stream));
  B200_CUDA_OK(cudaMemcpyAsync(d_polys, b.polys.data(), b.polys.size() * sizeof(RPoly), cudaMemcpyHostToDevice, stream));
  raster_kernel<<<b.n_tasks, 128, 0, stream>>>(d_edges, d_polys, static_cast<int>(b.polys.size()), b.n_tasks, b.w, index, d_layer);
  B200_CUDA_OK(cudaGetLastError());
  B200_CUDA_OK(cudaFreeAsync(d_edges, stream));
  B200_CUDA_OK(cudaFreeAsync(d_polys, stream));
  return 0;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200vit_raster_polygons(const double* h_xy, const int32_t* h_counts, int32_t n_polygons, int32_t h, int32_t w,
                                       uint8_t index, uint8_t* d_layer, b200vit_stream stream) {
  if (!h_xy || !h_counts || n_polygons < 0 || h <= 0 || w <= 0 || !d_layer) return fail(B200VIT_EINVAL, "raster_polygons: bad argument");
  int rc = check_arch();
  if (rc) return rc;
  Builder b;
  b.h = h, b.w = w;
  std::vector<int> ixy;
  size_t off = 0;
  for (int p = 0; p < n_polygons; ++p) {
    const int n = h_counts[p];
    if (n < 0) return fail(B200VIT_EINVAL, "raster_polygons: negative vertex count");
    ixy.resize(2 * static_cast<size_t>(n));
    for (int i = 0; i < 2 * n; ++i) {
      const double v = h_xy[off + i];
      if (!(std::fabs(v) < static_cast<double>(kCoordLimit))) return fail(B200VIT_EINVAL, "raster_polygons: coordinate out of range");
      ixy[i] = static_cast<int>(v);  // ImageDraw hands Pillow doubles; _draw_polygon truncates them
    }
    off += 2 * static_cast<size_t>(n);
    if (n < 2) continue;  // ImageDraw rejects them; nothing to fill
    if (!b.polygon(ixy.data(), n)) return fail(B200VIT_EINVAL, b.err);
  }
  return launch(b, index, d_layer, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int b200vit_raster_lines(const double* h_xy, int32_t n_points, int32_t width, int32_t h, int32_t w, uint8_t index,
                                    uint8_t* d_layer, b200vit_stream stream) {
  if (!h_xy || n_points < 0 || h <= 0 || w <= 0 || !d_layer) return fail(B200VIT_EINVAL, "raster_lines: bad argument");
  int rc = check_arch();
  if (rc) return rc;
  Builder b;
  b.h = h, b.w = w;
  for (int i = 1; i < n_points; ++i)
    if (!b.line(h_xy[2 * i - 2], h_xy[2 * i - 1], h_xy[2 * i], h_xy[2 * i + 1], width)) return fail(B200VIT_EINVAL, b.err);
  return launch(b, index, d_layer, reinterpret_cast<cudaStream_t>(stream));
}

// The reference's scribble path (:244-246): t = np.linspace(0, 1, n)[i]; the cubic evaluated left to right in float64
// with ** as libm pow, exactly as the Python expression does.
// numpy's float64 ** int is libm pow(x, (double)n); a compiler would turn pow(u, 2.0) into u * u, which differs from
// libm's result by one ulp in rare cases -- call through a volatile pointer so the library function really runs.
static double (*volatile libm_pow)(double, double) = static_cast<double (*)(double, double)>(std::pow);

extern "C" int b200vit_scribble_points(const double* h_ctrl8, int32_t n_points, double* h_xy_out) {
  if (!h_ctrl8 || !h_xy_out || n_points < 2) return fail(B200VIT_EINVAL, "scribble_points: bad argument");
  const double step = 1.0 / static_cast<double>(n_points - 1);
  for (int i = 0; i < n_points; ++i) {
    const double t = (i == n_points - 1) ? 1.0 : static_cast<double>(i) * step + 0.0;
    const double u = 1 - t;
    for (int c = 0; c < 2; ++c) {
      const double p0 = h_ctrl8[c], p1 = h_ctrl8[2 + c], p2 = h_ctrl8[4 + c], p3 = h_ctrl8[6 + c];
      h_xy_out[2 * i + c] = libm_pow(u, 3.0) * p0 + 3 * libm_pow(u, 2.0) * t * p1 + 3 * u * libm_pow(t, 2.0) * p2 + libm_pow(t, 3.0) * p3;
    }
  }
  return 0;
}
