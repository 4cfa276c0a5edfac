/* Plain-C caller of libb200vit.so: proves include/b200vit.h is a C header (no C++ types across the ABI) and shows
 * the call sequence another host language would bind (cgo / JNI / N-API / ctypes).
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -Lrga3-release_b200 -lb200vit -L/usr/local/cuda/lib64 -lcudart \
 *       -Wl,-rpath,$PWD/rga3-release_b200 -o /tmp/c_abi_demo
 *
 * Without a GPU it creates a plan on the host and prints its geometry.  With a GPU it also
 *   1. runs one b200vit_gemm (tcgen05 kernel, STORE_F32 epilogue) on integer-valued inputs and checks the result
 *      EXACTLY against a host loop,
 *   2. packs a tiny random tower from HOST float arrays with b200vit_pack_weights and runs b200vit_forward on a
 *      small grid from frames, checking the output is finite and reproducible bit for bit.
 * The CUDA runtime is only used for memory (cudaMalloc / cudaMemcpy); its five prototypes are declared here so the
 * file stays plain C99 without CUDA headers. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200vit.h"

extern int cudaGetDeviceCount(int* count);
extern int cudaMalloc(void** p, size_t bytes);
extern int cudaFree(void* p);
extern int cudaMemcpy(void* dst, const void* src, size_t bytes, int kind); /* 1 = H2D, 2 = D2H */
extern int cudaDeviceSynchronize(void);

static uint16_t bf16_of(float f) { /* exact for the small integers used here */
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)(u >> 16);
}

static void* dev_copy(const void* h, size_t bytes) {
  void* d = NULL;
  if (cudaMalloc(&d, bytes) != 0) return NULL;
  if (h != NULL && cudaMemcpy(d, h, bytes, 1) != 0) return NULL;
  return d;
}

static int check(int rc, const char* what) {
  if (rc != B200VIT_OK) fprintf(stderr, "%s failed (%d): %s\n", what, rc, b200vit_last_error());
  return rc;
}

static int demo_gemm(void) {
  enum { M = 256, N = 256, K = 128 };
  static uint16_t a[M * K], b[N * K];
  static float ia[M * K], ib[N * K], out[M * N];
  b200vit_gemm_args g;
  int i, j, k, bad = 0;
  for (i = 0; i < M * K; ++i) ia[i] = (float)((i * 7 + i / K) % 9 - 4), a[i] = bf16_of(ia[i]);
  for (i = 0; i < N * K; ++i) ib[i] = (float)((i * 5 + i / K) % 7 - 3), b[i] = bf16_of(ib[i]);
  memset(&g, 0, sizeof(g));
  g.d_a = dev_copy(a, sizeof(a)), g.d_b = dev_copy(b, sizeof(b)), g.d_out = dev_copy(NULL, sizeof(out));
  if (!g.d_a || !g.d_b || !g.d_out) return 1;
  g.m = M, g.n = N, g.k = K, g.ldo = N, g.epilogue = B200VIT_EPI_STORE_F32;
  if (check(b200vit_gemm(&g, NULL), "b200vit_gemm")) return 1;
  cudaDeviceSynchronize();
  cudaMemcpy(out, g.d_out, sizeof(out), 2);
  for (i = 0; i < M; ++i)
    for (j = 0; j < N; ++j) {
      float ref = 0.f;
      for (k = 0; k < K; ++k) ref += ia[i * K + k] * ib[j * K + k];
      if (ref != out[i * N + j]) ++bad;
    }
  printf("gemm %dx%dx%d on the GPU: %d mismatches against the host loop\n", M, N, K, bad);
  cudaFree((void*)g.d_a), cudaFree((void*)g.d_b), cudaFree(g.d_out);
  return bad != 0;
}

static float* rnd(size_t n, float scale, float offset, uint32_t* seed) {
  float* p = (float*)malloc(n * sizeof(float));
  size_t i;
  for (i = 0; i < n; ++i) {
    *seed = *seed * 1664525u + 1013904223u;
    p[i] = offset + scale * ((float)(*seed >> 8) / 8388608.0f - 1.0f);
  }
  return p;
}

static int demo_forward(void) {
  /* a 2-layer tower with head_dim 80, hidden 160: the smallest shape the kernels accept */
  enum { DEPTH = 2, D = 160, I = 200, O = 128, T = 4, H = 56, W = 84 };
  b200vit_cfg cfg;
  b200vit_raw_layer raw_layers[DEPTH];
  b200vit_raw_weights raw;
  b200vit_weights w;
  b200vit_layer_weights layers[DEPTH];
  b200vit_plan* plan = NULL;
  b200vit_frames fr;
  int64_t grid[3] = {T / 2, H / 14, W / 14};
  const size_t kpe = 3 * 2 * 14 * 14, m = (size_t)(T / 2) * (H / 14) * (W / 14);
  uint32_t seed = 1;
  size_t bytes, ws_bytes, i;
  void *d_packed, *d_ws, *d_frames, *d_out;
  uint8_t* frames = (uint8_t*)malloc((size_t)T * H * W * 3);
  float *o1 = (float*)malloc(m / 4 * O * sizeof(float)), *o2 = (float*)malloc(m / 4 * O * sizeof(float));
  int l, bad = 0;
  memset(&cfg, 0, sizeof(cfg));
  cfg.depth = DEPTH, cfg.hidden = D, cfg.intermediate = I, cfg.heads = 2, cfg.out_hidden = O;
  cfg.patch = 14, cfg.temporal_patch = 2, cfg.merge = 2, cfg.window = 112, cfg.in_channels = 3;
  cfg.n_fullatt = 1, cfg.fullatt[0] = 1;
  for (l = 0; l < DEPTH; ++l) {
    b200vit_raw_layer* r = &raw_layers[l];
    r->norm1_w = rnd(D, 0.1f, 1.f, &seed), r->norm2_w = rnd(D, 0.1f, 1.f, &seed);
    r->qkv_w = rnd(3 * D * D, 0.05f, 0.f, &seed), r->qkv_b = rnd(3 * D, 0.05f, 0.f, &seed);
    r->proj_w = rnd(D * D, 0.05f, 0.f, &seed), r->proj_b = rnd(D, 0.05f, 0.f, &seed);
    r->gate_w = rnd(I * D, 0.05f, 0.f, &seed), r->gate_b = rnd(I, 0.05f, 0.f, &seed);
    r->up_w = rnd(I * D, 0.05f, 0.f, &seed), r->up_b = rnd(I, 0.05f, 0.f, &seed);
    r->down_w = rnd(D * I, 0.05f, 0.f, &seed), r->down_b = rnd(D, 0.05f, 0.f, &seed);
  }
  raw.dtype = 0; /* fp32, host pointers */
  raw.patch_w = rnd(D * kpe, 0.05f, 0.f, &seed), raw.layers = raw_layers;
  raw.merger_ln_w = rnd(D, 0.1f, 1.f, &seed);
  raw.merger_fc1_w = rnd(16 * D * D, 0.05f, 0.f, &seed), raw.merger_fc1_b = rnd(4 * D, 0.05f, 0.f, &seed);
  raw.merger_fc2_w = rnd((size_t)O * 4 * D, 0.05f, 0.f, &seed), raw.merger_fc2_b = rnd(O, 0.05f, 0.f, &seed);
  bytes = b200vit_packed_weights_bytes(&cfg);
  d_packed = dev_copy(NULL, bytes);
  if (!d_packed || check(b200vit_pack_weights(&cfg, &raw, d_packed, bytes, &w, layers, NULL), "b200vit_pack_weights")) return 1;
  if (check(b200vit_plan_create(grid, 1, &cfg, &plan), "b200vit_plan_create")) return 1;
  ws_bytes = b200vit_workspace_bytes(plan);
  d_ws = dev_copy(NULL, ws_bytes + 1024);
  for (i = 0; i < (size_t)T * H * W * 3; ++i) frames[i] = (uint8_t)((i * 2654435761u) >> 24);
  d_frames = dev_copy(frames, (size_t)T * H * W * 3);
  d_out = dev_copy(NULL, m / 4 * O * sizeof(float));
  if (!d_ws || !d_frames || !d_out) return 1;
  fr.d_frames = (const uint8_t*)d_frames, fr.t = T, fr.h = H, fr.w = W;
  {
    void* ws_aligned = (void*)(((uintptr_t)d_ws + 1023) & ~(uintptr_t)1023);
    if (check(b200vit_forward(plan, &w, NULL, &fr, NULL, d_out, 1, NULL, ws_aligned, ws_bytes, NULL), "b200vit_forward")) return 1;
    cudaDeviceSynchronize();
    cudaMemcpy(o1, d_out, m / 4 * O * sizeof(float), 2);
    if (check(b200vit_forward(plan, &w, NULL, &fr, NULL, d_out, 1, NULL, ws_aligned, ws_bytes, NULL), "b200vit_forward")) return 1;
    cudaDeviceSynchronize();
    cudaMemcpy(o2, d_out, m / 4 * O * sizeof(float), 2);
  }
  for (i = 0; i < m / 4 * O; ++i)
    if (!isfinite(o1[i]) || memcmp(&o1[i], &o2[i], 4) != 0) ++bad;
  printf("forward from %d frames of %dx%d: %zu merged tokens x %d, out[0] = %.5f, %d non-finite or irreproducible values\n", T, H, W,
         m / 4, O, (double)o1[0], bad);
  b200vit_plan_destroy(plan);
  cudaFree(d_packed), cudaFree(d_ws), cudaFree(d_frames), cudaFree(d_out);
  return bad != 0;
}

int main(void) {
  b200vit_cfg cfg;
  b200vit_plan* plan = NULL;
  int64_t grid[3] = {8, 32, 32}; /* BASELINE config 2: 16 frames of 448x448 */
  int32_t full[4] = {7, 15, 23, 31};
  int i, n_gpus = 0;
  memset(&cfg, 0, sizeof(cfg));
  cfg.depth = 32, cfg.hidden = 1280, cfg.intermediate = 3420, cfg.heads = 16, cfg.out_hidden = 3584;
  cfg.patch = 14, cfg.temporal_patch = 2, cfg.merge = 2, cfg.window = 112, cfg.in_channels = 3;
  cfg.n_fullatt = 4;
  for (i = 0; i < 4; ++i) cfg.fullatt[i] = full[i];
  printf("b200vit version %d\n", b200vit_version());
  if (check(b200vit_plan_create(grid, 1, &cfg, &plan), "b200vit_plan_create")) return 1;
  printf("workspace bytes: %zu, launches per forward: %d, packed weights: %zu bytes\n", b200vit_workspace_bytes(plan),
         b200vit_forward_launches(plan, 1), b200vit_packed_weights_bytes(&cfg));
  {
    /* host-side views of the plan (window_index etc.) through the query entry point */
    int64_t n = b200vit_plan_get(plan, B200VIT_PLAN_WINDOW_INDEX, NULL, 0);
    printf("window_index holds %lld bytes\n", (long long)n);
  }
  b200vit_plan_destroy(plan);
  if (cudaGetDeviceCount(&n_gpus) != 0 || n_gpus <= 0) {
    printf("no GPU: device calls skipped\n");
    return 0;
  }
  if (demo_gemm()) return 2;
  if (demo_forward()) return 3;
  printf("device calls ok\n");
  return 0;
}
