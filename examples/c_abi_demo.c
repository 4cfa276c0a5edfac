/* Plain-C caller of libb200vit.so: proves include/b200vit.h is a C header (no C++ types across the ABI) and shows
 * the call sequence another host language would bind (cgo / JNI / N-API / ctypes).
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -Lrga3-release_b200 -lb200vit -Wl,-rpath,$PWD/rga3-release_b200 -o /tmp/c_abi_demo
 *
 * Run without arguments it only creates a plan on the host (no GPU needed) and prints its geometry; the device calls
 * are compiled (so every prototype is checked) but skipped unless a GPU buffer is supplied by a real host program. */
#include <stdio.h>
#include <stdlib.h>

#include "b200vit.h"

int main(void) {
  b200vit_cfg cfg;
  b200vit_plan* plan = NULL;
  int64_t grid[3] = {8, 32, 32}; /* BASELINE config 2: 16 frames of 448x448 */
  int32_t full[4] = {7, 15, 23, 31};
  int i;
  cfg.depth = 32, cfg.hidden = 1280, cfg.intermediate = 3420, cfg.heads = 16, cfg.out_hidden = 3584;
  cfg.patch = 14, cfg.temporal_patch = 2, cfg.merge = 2, cfg.window = 112, cfg.in_channels = 3;
  cfg.n_fullatt = 4;
  for (i = 0; i < 4; ++i) cfg.fullatt[i] = full[i];
  printf("b200vit version %d\n", b200vit_version());
  if (b200vit_plan_create(grid, 1, &cfg, &plan) != B200VIT_OK) {
    fprintf(stderr, "plan_create failed: %s\n", b200vit_last_error());
    return 1;
  }
  printf("workspace bytes: %zu, launches per forward: %d\n", b200vit_workspace_bytes(plan), b200vit_forward_launches(plan, 1));
  {
    /* host-side views of the plan (window_index etc.) through the query entry point */
    int64_t n = b200vit_plan_get(plan, 0, NULL, 0);
    printf("plan array 0 holds %lld bytes\n", (long long)n);
  }
  if (0) { /* device calls: prototypes checked at compile time, not executed here */
    b200vit_weights w;
    b200vit_frames fr;
    b200vit_overlay ov;
    b200vit_frame_op* d_ops = NULL;
    fr.d_frames = NULL, fr.t = 16, fr.h = 448, fr.w = 448;
    ov.kind = B200VIT_LAYER_NONE, ov.d_layer = NULL, ov.h_ops = NULL, ov.d_ops = d_ops, ov.d_ops_circle_r = -1;
    (void)b200vit_resize_bicubic(NULL, 16, 720, 1280, NULL, 392, 728, NULL, b200vit_resize_workspace_bytes(16, 720, 1280, 392, 728), NULL);
    (void)b200vit_stom_policy(NULL, NULL, 16, 1500, 8, 0, 448, 448, NULL, d_ops, NULL, b200vit_stom_policy_workspace_bytes(16, 1500, 448, 448), NULL);
    (void)b200vit_forward(plan, &w, NULL, &fr, &ov, NULL, 0, NULL, NULL, 0, NULL);
  }
  b200vit_plan_destroy(plan);
  return 0;
}
