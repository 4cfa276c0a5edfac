// C ABI: plan (everything derived from grid_thw), the whole-path forward and the
// single-op entry points.  See include/b200vit.h for the contract.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include <list>
#include <memory>

#include "internal.h"

namespace b200 {
struct CallCache {
  void* workspace = nullptr;
  std::vector<GemmPrepared> gemm;   // one memo per GEMM call site of forward()
  AttnPrepared attn;
  // optional per-launch profiling (cudaEvent pairs around every launch of a forward)
  std::vector<cudaEvent_t> ev;      // 2 per launch
  std::vector<int> ev_kind;         // B200VIT_K_* per launch
  size_t ev_used = 0;
  ~CallCache() {
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
  }
};
}  // namespace b200

using namespace b200;

struct b200vit_plan {
  b200vit_cfg cfg;
  std::vector<int64_t> grid;  // n*3
  int64_t m = 0;              // patches
  int unit = 4;
  int head_dim = 80;
  // host arrays
  std::vector<int64_t> window_index, reverse_index;
  std::vector<int32_t> cu_window, cu_full, row_map, merge_map, pos_ids;
  std::vector<float> rope_cos, rope_sin;  // [M, head_dim/2] in window order
  std::vector<float2> rope_table;         // [P, head_dim/4] fp32 (cos, sin) of coordinate * inv_freq: what the QKV epilogue reads
  std::vector<int32_t> rope_pos;          // [M, 2] (hpos, wpos) in window order
  std::vector<AttnTile> tiles_window, tiles_full;         // tcgen05 attention
  std::vector<int32_t> bounds_window, bounds_full;        // per-row [lo, hi) of the row's segment
  int maxblk_window = 0, maxblk_full = 0;
  bool uniform_windows = false;  // every window is 64 consecutive rows: the windowed layers' attention runs inside the QKV GEMM
  // workspace layout (byte offsets)
  size_t off_x = 0, off_h = 0, off_qkv = 0, off_attn = 0, off_act = 0, off_pv = 0, off_rowsq = 0, off_sync = 0, ws_bytes = 0;
  int ipad = 0, kpe = 0, rowsq_parts = 0;
  // device copies (lazy, then immutable: a plan is shared freely between threads and streams)
  mutable std::mutex mu;
  mutable bool uploaded = false;
  mutable int device = -1;
  mutable int32_t* d_row_map = nullptr;
  mutable int32_t* d_merge_map = nullptr;
  mutable float2* d_rope = nullptr;
  mutable int32_t* d_rope_pos = nullptr;
  mutable AttnTile* d_tiles_window = nullptr;
  mutable AttnTile* d_tiles_full = nullptr;
  mutable int32_t* d_bounds_window = nullptr;
  mutable int32_t* d_bounds_full = nullptr;
  // Host-side memos of one forward (tensor maps, launch geometry, profiling events) depend on the workspace and
  // weight pointers, not on the plan: one CallCache per workspace, looked up under `mu`, then used lock-free by the
  // calling thread (two concurrent forwards on the SAME workspace would race on the activations anyway).
  mutable std::list<std::unique_ptr<b200::CallCache>> caches;
  mutable b200::CallCache* last_cache = nullptr;
  mutable bool profile = false;
};

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// HF modeling_qwen2_5_vl.py:411-451 (get_window_index) + :476 (unique_consecutive)
void build_window_index(b200vit_plan& p) {
  const b200vit_cfg& c = p.cfg;
  const int win = c.window / c.merge / c.patch;
  int64_t base = 0;
  std::vector<int32_t> cu_raw{0};
  for (size_t gi = 0; gi < p.grid.size() / 3; ++gi) {
    const int64_t t = p.grid[gi * 3], h = p.grid[gi * 3 + 1], w = p.grid[gi * 3 + 2];
    const int64_t lh = h / c.merge, lw = w / c.merge;
    const int64_t pad_h = win - lh % win, pad_w = win - lw % win;  // a full extra window when divisible (:424-425)
    const int64_t nwh = (lh + pad_h) / win, nww = (lw + pad_w) / win;
    for (int64_t ti = 0; ti < t; ++ti)
      for (int64_t wh = 0; wh < nwh; ++wh)
        for (int64_t ww = 0; ww < nww; ++ww) {
          int cnt = 0;
          for (int ih = 0; ih < win; ++ih)
            for (int iw = 0; iw < win; ++iw) {
              const int64_t y = wh * win + ih, x = ww * win + iw;
              if (y < lh && x < lw) {
                p.window_index.push_back(base + ti * lh * lw + y * lw + x);
                ++cnt;
              }
            }
          cu_raw.push_back(cu_raw.back() + cnt * p.unit);
        }
    base += t * lh * lw;
  }
  p.cu_window.clear();
  for (int32_t v : cu_raw)
    if (p.cu_window.empty() || p.cu_window.back() != v) p.cu_window.push_back(v);
  p.reverse_index.assign(p.window_index.size(), 0);
  for (size_t i = 0; i < p.window_index.size(); ++i) p.reverse_index[p.window_index[i]] = static_cast<int64_t>(i);
}

// HF :382-409 (rot_pos_emb), :117-130 (inv_freq), :482-486 (reorder, cos/sin)
void build_rope(b200vit_plan& p) {
  const b200vit_cfg& c = p.cfg;
  const int half = p.head_dim / 2;   // rotary table width (40)
  const int nfreq = half / 2;        // 20 frequencies per axis
  std::vector<float> inv_freq(nfreq);
  for (int k = 0; k < nfreq; ++k) inv_freq[k] = 1.0f / powf(10000.0f, static_cast<float>(2 * k) / static_cast<float>(half));
  p.pos_ids.resize(p.m * 2);
  p.rope_cos.resize(p.m * half);
  p.rope_sin.resize(p.m * half);
  int64_t r = 0;
  const int mg = c.merge;
  for (size_t gi = 0; gi < p.grid.size() / 3; ++gi) {
    const int64_t t = p.grid[gi * 3], h = p.grid[gi * 3 + 1], w = p.grid[gi * 3 + 2];
    for (int64_t ti = 0; ti < t; ++ti)
      for (int64_t bh = 0; bh < h / mg; ++bh)
        for (int64_t bw = 0; bw < w / mg; ++bw)
          for (int ih = 0; ih < mg; ++ih)
            for (int iw = 0; iw < mg; ++iw, ++r) {
              const int hp = static_cast<int>(bh * mg + ih), wp = static_cast<int>(bw * mg + iw);
              p.pos_ids[r * 2] = hp;
              p.pos_ids[r * 2 + 1] = wp;
              const int64_t dst = p.row_map[r];
              for (int k = 0; k < nfreq; ++k) {
                const float ah = static_cast<float>(hp) * inv_freq[k], aw = static_cast<float>(wp) * inv_freq[k];
                p.rope_cos[dst * half + k] = cosf(ah);
                p.rope_sin[dst * half + k] = sinf(ah);
                p.rope_cos[dst * half + nfreq + k] = cosf(aw);
                p.rope_sin[dst * half + nfreq + k] = sinf(aw);
              }
            }
  }
  // the same values by coordinate (HF builds exactly this table, then gathers it with pos_ids, :396-408)
  int64_t side = 1;
  for (size_t gi = 0; gi < p.grid.size() / 3; ++gi) side = std::max(side, std::max(p.grid[gi * 3 + 1], p.grid[gi * 3 + 2]));
  p.rope_table.resize(side * nfreq);
  for (int64_t c0 = 0; c0 < side; ++c0)
    for (int k = 0; k < nfreq; ++k) {
      const float a = static_cast<float>(c0) * inv_freq[k];
      p.rope_table[c0 * nfreq + k] = make_float2(cosf(a), sinf(a));
    }
  p.rope_pos.resize(p.m * 2);
  for (int64_t r0 = 0; r0 < p.m; ++r0) {
    p.rope_pos[p.row_map[r0] * 2] = p.pos_ids[r0 * 2];
    p.rope_pos[p.row_map[r0] * 2 + 1] = p.pos_ids[r0 * 2 + 1];
  }
}

template <typename T>
int upload(T** dst, const std::vector<T>& src) {
  if (src.empty()) return 0;
  B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(dst), src.size() * sizeof(T)));
  B200_CUDA_OK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

int ensure_uploaded(const b200vit_plan* p) {
  std::lock_guard<std::mutex> lock(p->mu);
  int dev = 0;
  B200_CUDA_OK(cudaGetDevice(&dev));
  if (p->uploaded) {
    if (dev != p->device) return fail(B200VIT_EINVAL, "forward: this plan's tables live on another device (one plan per device)");
    return 0;
  }
  int rc;
  p->device = dev;
  if ((rc = upload(&p->d_row_map, p->row_map))) return rc;
  if ((rc = upload(&p->d_merge_map, p->merge_map))) return rc;
  if ((rc = upload(&p->d_rope, p->rope_table))) return rc;
  if ((rc = upload(&p->d_rope_pos, p->rope_pos))) return rc;
  if ((rc = upload(&p->d_tiles_window, p->tiles_window))) return rc;
  if ((rc = upload(&p->d_tiles_full, p->tiles_full))) return rc;
  if ((rc = upload(&p->d_bounds_window, p->bounds_window))) return rc;
  if ((rc = upload(&p->d_bounds_full, p->bounds_full))) return rc;
  p->uploaded = true;
  return 0;
}

// Records an event pair around one launch when profiling is on.
struct Prof {
  CallCache* p;
  cudaStream_t st;
  bool on;
  void begin(int kind) {
    if (!on) return;
    if (p->ev_used + 2 > p->ev.size()) {
      for (int i = 0; i < 2; ++i) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        p->ev.push_back(e);
      }
      p->ev_kind.push_back(kind);
    } else {
      p->ev_kind[p->ev_used / 2] = kind;
    }
    cudaEventRecord(p->ev[p->ev_used], st);
  }
  void end() {
    if (!on) return;
    cudaEventRecord(p->ev[p->ev_used + 1], st);
    p->ev_used += 2;
  }
};

// Keep the fp32 residual stream resident in L2 across the ~250 MB of other traffic each block generates:
// x is read by both RMSNorms and read-modify-written by both residual GEMMs of every block (SURVEY.md 8d:
// 252 MB of HBM traffic per block otherwise).  Implemented as a persisting access-policy window on the
// caller's stream for the duration of the forward.  B200VIT_L2_PERSIST=0 disables it.
// Policy (b200vit_set_l2_persist): 0 = never touch the device's persisting-L2 limit, 1 (default) = manage it.
// The limit is a per-DEVICE setting shared with the host application, so it is changed rarely and under a lock:
// grown (never shrunk) to the largest residual stream that fits, and given back only when a clip arrives whose
// residual stream does not fit at all.
struct L2State {
  bool probed = false;
  size_t max_persist = 0, max_window = 0, set_aside = 0;
};
std::mutex g_l2_mu;
L2State g_l2[64];
int g_l2_mode = -1;

struct L2Persist {
  cudaStream_t stream;
  bool active = false;
  L2Persist(cudaStream_t st, void* base, size_t bytes) : stream(st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || bytes == 0) return;
    size_t max_window = 0, max_persist = 0;
    {
      std::lock_guard<std::mutex> lock(g_l2_mu);
      if (g_l2_mode < 0) {
        const char* e = getenv("B200VIT_L2_PERSIST");
        g_l2_mode = (e == nullptr || e[0] != '0') ? 1 : 0;
      }
      if (g_l2_mode == 0) return;
      L2State& s = g_l2[dev];
      if (!s.probed) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, dev);
        s.max_persist = static_cast<size_t>(v > 0 ? v : 0);
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        s.max_window = static_cast<size_t>(v > 0 ? v : 0);
        s.probed = true;
        cudaGetLastError();
      }
      if (s.max_persist == 0) return;
      if (bytes > s.max_persist || bytes > s.max_window) {
        // long clips: the residual stream does not fit the persisting partition; a partial window only shrinks the
        // normal L2 (measured on cfg 4) -- give the carve-out back and stream
        if (s.set_aside != 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0) == cudaSuccess) s.set_aside = 0;
        cudaGetLastError();
        return;
      }
      if (bytes > s.set_aside) {
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes) != cudaSuccess) {
          cudaGetLastError();
          return;
        }
        s.set_aside = bytes;
      }
      max_window = s.max_window, max_persist = s.max_persist;
    }
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = base;
    attr.accessPolicyWindow.num_bytes = bytes;
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    (void)max_window, (void)max_persist;
    active = cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
    cudaGetLastError();
  }
  ~L2Persist() {
    if (!active) return;
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.num_bytes = 0;  // disables the window for later work on this stream
    cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
  }
};

bool is_fullatt(const b200vit_cfg& c, int layer) {
  for (int i = 0; i < c.n_fullatt; ++i)
    if (c.fullatt[i] == layer) return true;
  return false;
}

}  // namespace

extern "C" {

int b200vit_plan_create(const int64_t* h_grid_thw, int n_grids, const b200vit_cfg* cfg, b200vit_plan** out) {
  if (!h_grid_thw || n_grids <= 0 || !cfg || !out) return fail(B200VIT_EINVAL, "plan_create: null argument");
  const b200vit_cfg& c = *cfg;
  if (c.depth <= 0 || c.hidden <= 0 || c.heads <= 0 || c.hidden % c.heads) return fail(B200VIT_EINVAL, "plan_create: bad hidden/heads");
  if (c.hidden / c.heads != 80) return fail(B200VIT_EINVAL, "plan_create: kernels are specialised for head_dim 80 (all Qwen2.5-VL towers)");
  if (c.merge <= 0 || c.patch <= 0 || c.temporal_patch <= 0 || c.window % (c.merge * c.patch) || c.window / c.merge / c.patch <= 0)
    return fail(B200VIT_EINVAL, "plan_create: bad patch/merge/window");
  if (c.n_fullatt < 0 || c.n_fullatt > 64) return fail(B200VIT_EINVAL, "plan_create: bad fullatt list");
  if (c.hidden % 8 || c.out_hidden % 8 || c.intermediate <= 0) return fail(B200VIT_EINVAL, "plan_create: dims must be multiples of 8");
  if ((c.in_channels * c.temporal_patch * c.patch * c.patch) % 8) return fail(B200VIT_EINVAL, "plan_create: patch vector length must be a multiple of 8");
  b200vit_plan* p = new (std::nothrow) b200vit_plan();
  if (!p) return fail(B200VIT_ENOMEM, "plan_create: out of host memory");
  p->cfg = c;
  p->unit = c.merge * c.merge;
  p->head_dim = c.hidden / c.heads;
  p->grid.assign(h_grid_thw, h_grid_thw + 3 * n_grids);
  for (int i = 0; i < n_grids; ++i) {
    const int64_t t = h_grid_thw[i * 3], h = h_grid_thw[i * 3 + 1], w = h_grid_thw[i * 3 + 2];
    if (t <= 0 || h <= 0 || w <= 0 || h % c.merge || w % c.merge) {
      delete p;
      return fail(B200VIT_EINVAL, "plan_create: grid_thw entries must be positive with h, w multiples of merge size");
    }
    p->m += t * h * w;
  }
  if (p->m > (int64_t(1) << 30)) {
    delete p;
    return fail(B200VIT_EINVAL, "plan_create: too many patches");
  }
  build_window_index(*p);
  // HF :488-496
  p->cu_full.push_back(0);
  for (int i = 0; i < n_grids; ++i)
    for (int64_t ti = 0; ti < h_grid_thw[i * 3]; ++ti)
      p->cu_full.push_back(p->cu_full.back() + static_cast<int32_t>(h_grid_thw[i * 3 + 1] * h_grid_thw[i * 3 + 2]));
  // patch row r (group r/unit) lands at window-order row reverse[group]*unit + r%unit  (HF :478-481)
  p->row_map.resize(p->m);
  for (int64_t r = 0; r < p->m; ++r)
    p->row_map[r] = static_cast<int32_t>(p->reverse_index[r / p->unit] * p->unit + r % p->unit);
  // merged row i (window order) returns to original position window_index[i]  (HF :512-513)
  p->merge_map.resize(p->window_index.size());
  for (size_t i = 0; i < p->window_index.size(); ++i) p->merge_map[i] = static_cast<int32_t>(p->window_index[i]);
  build_rope(*p);
  build_attn_tiles(p->cu_window, static_cast<int>(p->m), 128, p->tiles_window, p->bounds_window);
  build_attn_tiles(p->cu_full, static_cast<int>(p->m), 256, p->tiles_full, p->bounds_full);
  p->uniform_windows = p->m % 64 == 0;
  for (size_t i = 0; i + 1 < p->cu_window.size() && p->uniform_windows; ++i)
    p->uniform_windows = p->cu_window[i + 1] - p->cu_window[i] == 64;
  for (const AttnTile& t : p->tiles_window) p->maxblk_window = std::max(p->maxblk_window, t.n_kv_blocks);
  for (const AttnTile& t : p->tiles_full) p->maxblk_full = std::max(p->maxblk_full, t.n_kv_blocks);
  // workspace
  p->ipad = static_cast<int>(align_up(c.intermediate, 128));
  p->kpe = c.in_channels * c.temporal_patch * c.patch * c.patch;
  const size_t M = static_cast<size_t>(p->m), D = c.hidden;
  size_t off = 0;
  p->off_x = off, off = align_up(off + M * D * 4, 1024);
  p->off_h = off, off = align_up(off + M * D * 2, 1024);
  p->off_qkv = off, off = align_up(off + M * 3 * D * 2, 1024);
  p->off_attn = off, off = align_up(off + M * D * 2, 1024);
  p->off_act = off, off = align_up(off + M * p->ipad * 2, 1024);
  p->off_pv = off, off = align_up(off + M * p->kpe * 2, 1024);
  p->rowsq_parts = (c.hidden + 127) / 128;
  p->off_rowsq = off, off = align_up(off + M * p->rowsq_parts * 4, 1024);
  p->off_sync = off, off = align_up(off + B200VIT_GEMM_SYNC_INTS * 4, 1024);
  p->ws_bytes = off;
  *out = p;
  return 0;
}

void b200vit_plan_destroy(b200vit_plan* p) {
  if (!p) return;
  cudaFree(p->d_row_map);
  cudaFree(p->d_merge_map);
  cudaFree(p->d_rope);
  cudaFree(p->d_rope_pos);
  cudaFree(p->d_tiles_window);
  cudaFree(p->d_tiles_full);
  cudaFree(p->d_bounds_window);
  cudaFree(p->d_bounds_full);
  delete p;
}

int64_t b200vit_plan_get(const b200vit_plan* p, int which, void* h_dst, size_t cap) {
  if (!p) return fail(B200VIT_EINVAL, "plan_get: null plan");
  const void* src = nullptr;
  size_t bytes = 0;
  switch (which) {
    case B200VIT_PLAN_M: src = &p->m, bytes = sizeof(int64_t); break;
    case B200VIT_PLAN_WINDOW_INDEX: src = p->window_index.data(), bytes = p->window_index.size() * 8; break;
    case B200VIT_PLAN_REVERSE_INDEX: src = p->reverse_index.data(), bytes = p->reverse_index.size() * 8; break;
    case B200VIT_PLAN_CU_WINDOW: src = p->cu_window.data(), bytes = p->cu_window.size() * 4; break;
    case B200VIT_PLAN_CU_FULL: src = p->cu_full.data(), bytes = p->cu_full.size() * 4; break;
    case B200VIT_PLAN_ROW_MAP: src = p->row_map.data(), bytes = p->row_map.size() * 4; break;
    case B200VIT_PLAN_ROPE_COS: src = p->rope_cos.data(), bytes = p->rope_cos.size() * 4; break;
    case B200VIT_PLAN_ROPE_SIN: src = p->rope_sin.data(), bytes = p->rope_sin.size() * 4; break;
    case B200VIT_PLAN_POS_IDS: src = p->pos_ids.data(), bytes = p->pos_ids.size() * 4; break;
    case B200VIT_PLAN_ROPE_TABLE: src = p->rope_table.data(), bytes = p->rope_table.size() * 8; break;
    case B200VIT_PLAN_ROPE_POS: src = p->rope_pos.data(), bytes = p->rope_pos.size() * 4; break;
    default: return fail(B200VIT_EINVAL, "plan_get: unknown array id");
  }
  if (h_dst && cap >= bytes) std::memcpy(h_dst, src, bytes);
  return static_cast<int64_t>(bytes);
}

size_t b200vit_workspace_bytes(const b200vit_plan* p) { return p ? p->ws_bytes : 0; }

static bool fuse_window_attention(const b200vit_plan* p) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200VIT_FUSE_WINATTN");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1 && p->uniform_windows && p->cfg.window / p->cfg.merge / p->cfg.patch == 4 && p->unit == 4;
}

int b200vit_forward_launches(const b200vit_plan* p, int with_frames) {
  if (!p) return 0;
  const int t_pad_windows = 1;  // one overlay/patchify launch per 128 frames; clips here are <= 128 frames
  int n_full = 0;
  for (int l = 0; l < p->cfg.depth; ++l) n_full += is_fullatt(p->cfg, l) ? 1 : 0;
  const int fused = fuse_window_attention(p) ? p->cfg.depth - n_full : 0;   // those layers have no attention launch
  return (with_frames ? t_pad_windows : 0) + 1 + p->cfg.depth * 5 - fused + 3;
}

// The per-workspace memo of this plan (created on first use; at most 8 kept, oldest dropped first).
static CallCache* call_cache(const b200vit_plan* p, void* workspace, size_t gemm_sites) {
  std::lock_guard<std::mutex> lock(p->mu);
  for (auto it = p->caches.begin(); it != p->caches.end(); ++it)
    if ((*it)->workspace == workspace) {
      if (it != p->caches.begin()) p->caches.splice(p->caches.begin(), p->caches, it);
      p->last_cache = p->caches.front().get();
      return p->last_cache;
    }
  if (p->caches.size() >= 8) p->caches.pop_back();
  p->caches.emplace_front(new CallCache());
  CallCache* c = p->caches.front().get();
  c->workspace = workspace;
  c->gemm.assign(gemm_sites, GemmPrepared());
  p->last_cache = c;
  return c;
}

int b200vit_forward(const b200vit_plan* p, const b200vit_weights* w, const void* d_pixel_values, const b200vit_frames* frames,
                    const b200vit_overlay* overlay, void* d_out, int out_f32, float* d_last_hidden, void* d_workspace,
                    size_t workspace_bytes, b200vit_stream stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!p || !w || !d_out || !d_workspace) return fail(B200VIT_EINVAL, "forward: null argument");
  if ((d_pixel_values == nullptr) == (frames == nullptr))
    return fail(B200VIT_EINVAL, "forward: pass exactly one of d_pixel_values and frames");
  if (workspace_bytes < p->ws_bytes) return fail(B200VIT_ENOMEM, "forward: workspace smaller than b200vit_workspace_bytes()");
  if (reinterpret_cast<uintptr_t>(d_workspace) & 1023) return fail(B200VIT_EALIGN, "forward: workspace must be 1024-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  if ((rc = ensure_uploaded(p))) return rc;
  const b200vit_cfg& c = p->cfg;
  if (w->ipad != p->ipad) return fail(B200VIT_EINVAL, "forward: weights packed with a different intermediate padding");
  const int M = static_cast<int>(p->m), D = c.hidden;
  uint8_t* ws = reinterpret_cast<uint8_t*>(d_workspace);
  float* x = reinterpret_cast<float*>(ws + p->off_x);          // fp32 residual stream, window order
  void* xb = ws + p->off_h;                                    // its bf16 copy = A operand of the QKV / gate-up GEMMs
  void* qkv = ws + p->off_qkv;
  void* attn = ws + p->off_attn;
  void* act = ws + p->off_act;
  void* pv = ws + p->off_pv;
  float* rowsq = reinterpret_cast<float*>(ws + p->off_rowsq);  // [parts][M] partial row sums of x^2
  int32_t* sync = reinterpret_cast<int32_t*>(ws + p->off_sync);

  CallCache* cc = call_cache(p, d_workspace, static_cast<size_t>(3 + 4 * c.depth));
  const bool fuse_win = fuse_window_attention(p);
  int site = 0;
  Prof prof{cc, stream, p->profile};
  cc->ev_used = 0;
  L2Persist keep_x(stream, x, static_cast<size_t>(M) * D * 4);
  // stream-K hand-over flags: every launch leaves them zeroed, but the workspace is the caller's (first use, or a
  // forward that failed half-way)
  B200_CUDA_OK(cudaMemsetAsync(sync, 0, B200VIT_GEMM_SYNC_INTS * 4, stream));
  const void* a0 = d_pixel_values;
  if (frames) {
    if (p->grid.size() != 3) return fail(B200VIT_EINVAL, "forward: the frames entry takes a single-clip plan");
    const int64_t t_need = p->grid[0] * c.temporal_patch;
    if (frames->h != p->grid[1] * c.patch || frames->w != p->grid[2] * c.patch ||
        (frames->t != t_need && frames->t != t_need - (c.temporal_patch - 1)))
      return fail(B200VIT_EINVAL, "forward: frames shape does not match the plan's grid_thw");
    prof.begin(B200VIT_K_OVERLAY_PATCHIFY);
    if ((rc = launch_overlay_patchify(*frames, overlay, c.patch, c.temporal_patch, c.merge, pv, nullptr, stream))) return rc;
    prof.end();
    a0 = pv;
  }
  b200vit_gemm_args g;
  auto gemm = [&](int kind) {
    prof.begin(kind);
    const int r = launch_gemm(g, stream, &cc->gemm[site++]);
    prof.end();
    return r;
  };
  // patch embed (+ window reorder in the store); also the first RMSNorm's inputs: bf16(x) and the row sums of x^2
  std::memset(&g, 0, sizeof(g));
  g.d_a = a0, g.d_b = w->patch_w, g.d_out = x, g.d_row_map = p->d_row_map;
  g.d_out_bf16 = xb, g.d_rowsq_out = rowsq;
  g.m = M, g.n = D, g.k = p->kpe, g.ldo = D, g.epilogue = B200VIT_EPI_STORE_F32;
  if ((rc = gemm(B200VIT_K_PATCH_EMBED))) return rc;

  int n_full_seen = 0;
  for (int l = 0; l < c.depth; ++l) {
    const b200vit_layer_weights& lw = w->layers[l];
    const bool full = is_fullatt(c, l);
    // qkv = rope(rstd(x) * (bf16(x) (Wqkv diag(gamma1))^T) + b)          (norm1 + qkv + RoPE, HF :313-316, :231-241)
    std::memset(&g, 0, sizeof(g));
    g.d_a = xb, g.d_b = lw.qkv_w, g.d_out = qkv, g.d_bias = lw.qkv_b, g.d_rope = p->d_rope, g.d_rope_pos = p->d_rope_pos;
    g.d_rowsq_in = rowsq, g.rowsq_parts = p->rowsq_parts, g.norm_eps = 1e-6f;
    g.m = M, g.n = 3 * D, g.k = D, g.ldo = 3 * D, g.epilogue = B200VIT_EPI_QKV_ROPE;
    if (!full && fuse_win) {
      // windowed layer, every window = 64 consecutive rows: the attention runs inside the QKV epilogue (HF :244-283)
      g.d_out = attn, g.ldo = D, g.epilogue = B200VIT_EPI_QKV_ROPE_WINATTN;
      if ((rc = gemm(B200VIT_K_QKV_WINATTN))) return rc;
    } else {
      if ((rc = gemm(B200VIT_K_QKV))) return rc;
      prof.begin(full ? B200VIT_K_ATTN_FULL : B200VIT_K_ATTN_WINDOW);
      // full layers draw their work items from a counter in the tail of the (zeroed) sync scratch, one per layer
      int32_t* counter = (full && n_full_seen < 64) ? sync + (B200VIT_GEMM_SYNC_INTS - 64) + n_full_seen : nullptr;
      if (full) ++n_full_seen;
      rc = launch_attention_tc(qkv, attn, full ? p->d_tiles_full : p->d_tiles_window,
                               static_cast<int>(full ? p->tiles_full.size() : p->tiles_window.size()), full ? 256 : 128,
                               full ? p->maxblk_full : p->maxblk_window, full ? p->d_bounds_full : p->d_bounds_window, M,
                               c.heads, stream, &cc->attn, counter);
      if (rc) return rc;
      prof.end();
    }
    // x += attn Wp^T + b; xb = bf16(x); rowsq = partial row sums of x^2                            (HF :313-316)
    std::memset(&g, 0, sizeof(g));
    g.d_a = attn, g.d_b = lw.proj_w, g.d_out = x, g.d_bias = lw.proj_b;
    g.d_out_bf16 = xb, g.d_rowsq_out = rowsq, g.d_sync = sync;
    g.m = M, g.n = D, g.k = D, g.ldo = D, g.epilogue = B200VIT_EPI_BIAS_RESIDUAL_NORM;
    if ((rc = gemm(B200VIT_K_PROJ))) return rc;
    // act = silu(g) * u with [g u] = rstd(x) * (bf16(x) (Wgu diag(gamma2))^T) + b            (norm2 + gate/up, HF :88)
    std::memset(&g, 0, sizeof(g));
    g.d_a = xb, g.d_b = lw.gateup_w, g.d_out = act, g.d_bias = lw.gateup_b;
    g.d_rowsq_in = rowsq, g.rowsq_parts = p->rowsq_parts, g.norm_eps = 1e-6f;
    g.m = M, g.n = 2 * p->ipad, g.k = D, g.ldo = p->ipad, g.epilogue = B200VIT_EPI_SWIGLU;
    if ((rc = gemm(B200VIT_K_GATEUP))) return rc;
    std::memset(&g, 0, sizeof(g));
    g.d_a = act, g.d_b = lw.down_w, g.d_out = x, g.d_bias = lw.down_b;
    g.d_out_bf16 = xb, g.d_rowsq_out = rowsq, g.d_sync = sync;
    g.m = M, g.n = D, g.k = p->ipad, g.ldo = D, g.epilogue = B200VIT_EPI_BIAS_RESIDUAL_NORM;
    if ((rc = gemm(B200VIT_K_DOWN))) return rc;
  }
  if (d_last_hidden)
    B200_CUDA_OK(cudaMemcpyAsync(d_last_hidden, x, static_cast<size_t>(M) * D * 4, cudaMemcpyDeviceToDevice, stream));
  // merger (HF :133-146) + un-reorder (:512-513).  ln_q is the one RMSNorm left as a kernel: its output is viewed
  // [M/4, 4D], so four different row scales meet inside one K reduction of fc1.
  const int Mm = M / p->unit, Dm = D * p->unit;
  prof.begin(B200VIT_K_RMSNORM);
  if ((rc = launch_rmsnorm(x, w->merger_ln_w, xb, M, D, 1e-6f, stream))) return rc;
  prof.end();
  std::memset(&g, 0, sizeof(g));
  g.d_a = xb, g.d_b = w->merger_fc1_w, g.d_out = attn, g.d_bias = w->merger_fc1_b;
  g.m = Mm, g.n = Dm, g.k = Dm, g.ldo = Dm, g.epilogue = B200VIT_EPI_BIAS_GELU;
  if ((rc = gemm(B200VIT_K_MERGER_FC1))) return rc;
  std::memset(&g, 0, sizeof(g));
  g.d_a = attn, g.d_b = w->merger_fc2_w, g.d_out = d_out, g.d_bias = w->merger_fc2_b, g.d_row_map = p->d_merge_map;
  g.m = Mm, g.n = c.out_hidden, g.k = Dm, g.ldo = c.out_hidden;
  g.epilogue = out_f32 ? B200VIT_EPI_BIAS_F32 : B200VIT_EPI_BIAS_BF16;
  if ((rc = gemm(B200VIT_K_MERGER_FC2))) return rc;
  return 0;
}

int b200vit_set_l2_persist(int mode) {
  if (mode != 0 && mode != 1) return fail(B200VIT_EINVAL, "set_l2_persist: mode must be 0 or 1");
  std::lock_guard<std::mutex> lock(g_l2_mu);
  g_l2_mode = mode;
  return 0;
}

int b200vit_profile_enable(b200vit_plan* p, int enable) {
  if (!p) return fail(B200VIT_EINVAL, "profile_enable: null plan");
  std::lock_guard<std::mutex> lock(p->mu);
  p->profile = enable != 0;
  return 0;
}

int b200vit_profile_read(b200vit_plan* p, float* h_ms_by_kind, int32_t* h_count_by_kind) {
  if (!p || !h_ms_by_kind || !h_count_by_kind) return fail(B200VIT_EINVAL, "profile_read: null argument");
  for (int i = 0; i < B200VIT_K_COUNT; ++i) h_ms_by_kind[i] = 0.f, h_count_by_kind[i] = 0;
  CallCache* cc;
  {
    std::lock_guard<std::mutex> lock(p->mu);
    cc = p->last_cache;
  }
  if (!cc) return 0;
  for (size_t i = 0; i + 1 < cc->ev_used; i += 2) {
    B200_CUDA_OK(cudaEventSynchronize(cc->ev[i + 1]));
    float ms = 0.f;
    B200_CUDA_OK(cudaEventElapsedTime(&ms, cc->ev[i], cc->ev[i + 1]));
    const int k = cc->ev_kind[i / 2];
    h_ms_by_kind[k] += ms;
    h_count_by_kind[k] += 1;
  }
  return 0;
}

// ---------------------------------------------------------------- single ops
int b200vit_overlay_composite(const b200vit_frames* frames, const b200vit_overlay* overlay, uint8_t* d_out,
                              b200vit_stream stream) {
  if (!frames || !d_out) return fail(B200VIT_EINVAL, "overlay_composite: null argument");
  int rc = check_arch();
  if (rc) return rc;
  return launch_overlay_patchify(*frames, overlay, 14, 2, 2, nullptr, d_out, reinterpret_cast<cudaStream_t>(stream));
}

int b200vit_overlay_patchify(const b200vit_frames* frames, const b200vit_overlay* overlay, int patch, int tps, int merge,
                             void* d_out_bf16, b200vit_stream stream) {
  if (!frames || !d_out_bf16) return fail(B200VIT_EINVAL, "overlay_patchify: null argument");
  int rc = check_arch();
  if (rc) return rc;
  return launch_overlay_patchify(*frames, overlay, patch, tps, merge, d_out_bf16, nullptr, reinterpret_cast<cudaStream_t>(stream));
}

int b200vit_gemm(const b200vit_gemm_args* args, b200vit_stream stream) {
  if (!args) return fail(B200VIT_EINVAL, "gemm: null argument");
  int rc = check_arch();
  if (rc) return rc;
  return launch_gemm(*args, reinterpret_cast<cudaStream_t>(stream));
}

int b200vit_rmsnorm(const float* d_x, const float* d_w, void* d_out_bf16, int rows, int dim, float eps, b200vit_stream stream) {
  int rc = check_arch();
  if (rc) return rc;
  return launch_rmsnorm(d_x, d_w, d_out_bf16, rows, dim, eps, reinterpret_cast<cudaStream_t>(stream));
}

int b200vit_attention(const void* d_qkv, void* d_out, const int32_t* h_cu_seqlens, int n_segments, int heads,
                      b200vit_stream stream_) {
  if (!d_qkv || !d_out || !h_cu_seqlens || n_segments < 0 || heads <= 0) return fail(B200VIT_EINVAL, "attention: bad argument");
  int rc = check_arch();
  if (rc) return rc;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  std::vector<int32_t> cu(h_cu_seqlens, h_cu_seqlens + n_segments + 1);
  const int m_rows = cu.back();
  std::vector<AttnTile> tiles;
  std::vector<int32_t> bounds;
  int max_len = 0;
  for (int i = 0; i < n_segments; ++i) max_len = std::max(max_len, cu[i + 1] - cu[i]);
  const int rows_per_tile = max_len > 256 ? 256 : 128;  // long segments: two query tiles per CTA share K/V blocks
  build_attn_tiles(cu, m_rows, rows_per_tile, tiles, bounds);
  if (tiles.empty()) return 0;
  int maxblk = 0;
  for (const AttnTile& t : tiles) maxblk = std::max(maxblk, t.n_kv_blocks);
  AttnTile* d_tiles = nullptr;
  int32_t* d_bounds = nullptr;
  B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&d_tiles), tiles.size() * sizeof(AttnTile)));
  B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&d_bounds), bounds.size() * sizeof(int32_t)));
  B200_CUDA_OK(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * sizeof(AttnTile), cudaMemcpyHostToDevice, stream));
  B200_CUDA_OK(cudaMemcpyAsync(d_bounds, bounds.data(), bounds.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  int32_t* d_counter = nullptr;  // full-attention kernel: work items drawn from a zeroed counter, as in b200vit_forward
  if (rows_per_tile == 256) {
    B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&d_counter), sizeof(int32_t)));
    B200_CUDA_OK(cudaMemsetAsync(d_counter, 0, sizeof(int32_t), stream));
  }
  rc = launch_attention_tc(d_qkv, d_out, d_tiles, static_cast<int>(tiles.size()), rows_per_tile, maxblk, d_bounds, m_rows, heads,
                           stream, nullptr, d_counter);
  cudaStreamSynchronize(stream);  // test entry point: the temporaries are freed before returning
  cudaFree(d_tiles);
  cudaFree(d_bounds);
  if (d_counter) cudaFree(d_counter);
  return rc;
}

int b200vit_cast_to_bf16(const void* d_in, int in_dtype, void* d_out, int64_t n, b200vit_stream stream) {
  int rc = check_arch();
  if (rc) return rc;
  return launch_cast_bf16(d_in, in_dtype, d_out, n, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
