/* b200vit.h -- C ABI of the B200-native RGA3 visual path.
 *
 * One shared library (libb200vit.so, sm_100a) replaces, behind the reference's
 * own module boundary, everything between "uint8 frames + visual-prompt layer"
 * and "merged visual embeddings":
 *
 *   STOM overlay      /root/reference/model/STOM.py:72-207 (warp :145-160, warp_point :163-207)
 *                     /root/reference/utils/visual_prompt_generator.py:102-104, :284-363
 *   normalise+patchify HF transformers video_processing_qwen2_vl.py:239-272,
 *                     image_processing_backends.py:291-331 (called from
 *                     /root/reference/utils/dataset.py:77-84)
 *   vision tower      HF transformers modeling_qwen2_5_vl.py:455-518
 *                     (Qwen2_5_VisionTransformerPretrainedModel.forward), called by the
 *                     reference at /root/reference/model/qwen_2_5_vl_sam2.py:182-200, :346-355
 *
 * Conventions: plain C, no C++/torch types; every pointer named d_* is a DEVICE
 * pointer owned by the caller; h_* is a HOST pointer.  All work is enqueued on
 * the caller's stream, nothing synchronises the host except where stated.
 * Return 0 on success, a negative B200VIT_E* code otherwise (never aborts, never
 * falls back to a CPU path); b200vit_last_error() gives the thread-local message.
 */
#ifndef B200VIT_H_
#define B200VIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VIT_VERSION 3

enum {
  B200VIT_OK = 0,
  B200VIT_EINVAL = -1,   /* bad shape / argument                         */
  B200VIT_EALIGN = -2,   /* pointer or leading dimension not 16-B aligned */
  B200VIT_EARCH = -3,    /* device is not sm_100                          */
  B200VIT_ECUDA = -4,    /* CUDA runtime / driver error                   */
  B200VIT_ENOMEM = -5    /* workspace too small                           */
};

typedef void* b200vit_stream; /* cudaStream_t */

/* ------------------------------------------------------------------ config
 * Mirrors Qwen2_5_VLVisionConfig (HF configuration_qwen2_5_vl.py:37-63).      */
typedef struct b200vit_cfg {
  int32_t depth, hidden, intermediate, heads, out_hidden;
  int32_t patch, temporal_patch, merge, window, in_channels;
  int32_t n_fullatt;
  int32_t fullatt[64];
} b200vit_cfg;

/* ------------------------------------------------------------------ plan
 * Everything that depends only on grid_thw: window_index / reverse index
 * (HF modeling :411-451, :512), cu_seqlens (:488-496), cu_window_seqlens
 * (:476), rope cos/sin tables in window order (:382-409, :485-486), attention
 * work lists, workspace layout.  Host part is computed at create time (no GPU
 * needed); device copies are uploaded lazily by the first forward (to the device
 * current at that moment: one plan per device).  After that a plan is immutable
 * and may be shared by any number of threads and streams: b200vit_forward takes
 * it const, and everything a call mutates lives in the caller's workspace (the
 * host-side memo of tensor maps per workspace is internal and locked).        */
typedef struct b200vit_plan b200vit_plan;

int b200vit_version(void);
const char* b200vit_last_error(void);

int b200vit_plan_create(const int64_t* h_grid_thw, int n_grids, const b200vit_cfg* cfg, b200vit_plan** out);
void b200vit_plan_destroy(b200vit_plan* plan);

enum {
  B200VIT_PLAN_M = 0,              /* int64 [1]   number of patches                      */
  B200VIT_PLAN_WINDOW_INDEX = 1,   /* int64 [M/4] HF window_index                        */
  B200VIT_PLAN_REVERSE_INDEX = 2,  /* int64 [M/4] argsort(window_index)                  */
  B200VIT_PLAN_CU_WINDOW = 3,      /* int32 [...] cu_window_seqlens after unique_consecutive */
  B200VIT_PLAN_CU_FULL = 4,        /* int32 [...] cu_seqlens                             */
  B200VIT_PLAN_ROW_MAP = 5,        /* int32 [M]   patch row -> row in window order       */
  B200VIT_PLAN_ROPE_COS = 6,       /* fp32  [M,head_dim/2] window order                  */
  B200VIT_PLAN_ROPE_SIN = 7,       /* fp32  [M,head_dim/2] window order                  */
  B200VIT_PLAN_POS_IDS = 8,        /* int32 [M,2] (hpos,wpos), ORIGINAL patch order       */
  B200VIT_PLAN_ROPE_TABLE = 9,     /* fp32 [P,head_dim/4,2] (cos,sin) by coordinate, P = largest grid side: what the QKV epilogue reads */
  B200VIT_PLAN_ROPE_POS = 10       /* int32 [M,2] (hpos,wpos), WINDOW order: the row positions the QKV epilogue reads   */
};
/* Copies a host-side plan array into h_dst (cap bytes).  Returns the number of
 * bytes the array holds (so a call with cap = 0 sizes it), negative on error.  */
int64_t b200vit_plan_get(const b200vit_plan* plan, int which, void* h_dst, size_t cap);
size_t b200vit_workspace_bytes(const b200vit_plan* plan);

/* ------------------------------------------------------------------ weights
 * Packed once from the HF state_dict by b200vit_pack_weights (below); all bf16
 * matrices are [N, K] row-major (nn.Linear layout), K padded as noted.  The
 * per-block RMSNorms (HF modeling :57-71, :313-320) do not exist as kernels:
 * RMSNorm(x) W^T == rstd(x) * (x (W diag(gamma))^T), so gamma is folded into the
 * columns of the following weight matrix and rstd is applied per row in that
 * GEMM's epilogue (SURVEY.md 2.1 "RMSNorm x65").                               */
typedef struct b200vit_layer_weights {
  const void* qkv_w;        /* bf16 [3D, D], rows head-major [(head, {q,k,v}, head_dim)], column k scaled by norm1.weight[k] */
  const float* qkv_b;       /* [3D], same row order                                     */
  const void* proj_w;       /* bf16 [D, D]                                              */
  const float* proj_b;      /* [D]                                                      */
  const void* gateup_w;     /* bf16 [2*Ipad, D], rows interleaved g0,u0,g1,u1,...,       */
                            /* column k scaled by norm2.weight[k]                        */
  const float* gateup_b;    /* [2*Ipad] interleaved the same way                         */
  const void* down_w;       /* bf16 [D, Ipad] (zero columns beyond I)                    */
  const float* down_b;      /* [D]                                                      */
} b200vit_layer_weights;

typedef struct b200vit_weights {
  const void* patch_w;      /* bf16 [D, C*tp*p*p] (Conv3d weight viewed 2-D, HF :106-114) */
  const b200vit_layer_weights* layers; /* HOST array [depth] of device pointers          */
  const float* merger_ln_w; /* [D]  (the merger's ln_q stays a kernel: its output is viewed [M/4, 4D]) */
  const void* merger_fc1_w; /* bf16 [4D, 4D]                                            */
  const float* merger_fc1_b;
  const void* merger_fc2_w; /* bf16 [out_hidden, 4D]                                    */
  const float* merger_fc2_b;
  int32_t ipad;             /* padded intermediate size (multiple of 128)               */
} b200vit_weights;

/* The tower's parameters as the HF state_dict holds them (names in SURVEY.md 8b):
 * every tensor row-major contiguous, all of one element type, HOST or DEVICE
 * pointers (detected per pointer).                                            */
typedef struct b200vit_raw_layer {
  const void *norm1_w, *qkv_w, *qkv_b, *proj_w, *proj_b;            /* [D] [3D,D] [3D] [D,D] [D] */
  const void *norm2_w, *gate_w, *gate_b, *up_w, *up_b, *down_w, *down_b; /* [D] [I,D] [I] [I,D] [I] [D,I] [D] */
} b200vit_raw_layer;
typedef struct b200vit_raw_weights {
  int32_t dtype;                       /* 0 fp32, 1 fp16, 2 bf16                         */
  const void* patch_w;                 /* [D, C, tp, p, p]                               */
  const b200vit_raw_layer* layers;     /* HOST array [depth]                             */
  const void *merger_ln_w, *merger_fc1_w, *merger_fc1_b, *merger_fc2_w, *merger_fc2_b;
} b200vit_raw_weights;

/* Packs the raw parameters into one caller-owned DEVICE buffer (256-byte aligned,
 * b200vit_packed_weights_bytes(cfg) bytes) and fills `out` / `out_layers`
 * (HOST array [depth]; out->layers points at it) with pointers into that buffer:
 * bf16 casts, gamma fold of norm1/norm2 into qkv/gate/up, head-major qkv rows, gate/up row
 * interleave, I -> Ipad zero padding.  Work is enqueued on `stream`; host-resident inputs are
 * staged synchronously.                                                        */
size_t b200vit_packed_weights_bytes(const b200vit_cfg* cfg);
int b200vit_pack_weights(const b200vit_cfg* cfg, const b200vit_raw_weights* raw, void* d_packed, size_t packed_bytes,
                         b200vit_weights* out, b200vit_layer_weights* out_layers, b200vit_stream stream);

/* ------------------------------------------------------------------ overlay
 * Output of the STOM policy (/root/reference/model/STOM.py:72-141), which stays
 * on the host: which frames get the prompt layer and how it is moved.         */
enum { B200VIT_LAYER_NONE = 0, B200VIT_LAYER_RGBA = 1, B200VIT_LAYER_PALETTE = 2, B200VIT_LAYER_BOX = 3 };
enum { B200VIT_FRAME_NONE = 0, B200VIT_FRAME_LAYER = 1, B200VIT_FRAME_CIRCLE = 2 };

typedef struct b200vit_frame_op {
  int32_t mode;        /* B200VIT_FRAME_*                                             */
  int32_t sx, sy;      /* integer translation of the layer (STOM.warp, :145-155)      */
  int32_t zx, zy;      /* 1: extra truncation-toward-zero source for column/row 0     */
  int32_t cx, cy, r;   /* circle stamp of STOM.warp_point (:195-201)                  */
  uint8_t rgba[4];     /* circle colour (alpha already clamped, :174)                  */
} b200vit_frame_op;

typedef struct b200vit_overlay {
  int32_t kind;                 /* B200VIT_LAYER_*                                     */
  const uint8_t* d_layer;       /* RGBA: [H,W,4]; PALETTE: [H,W] indices, 0 = clear     */
  uint8_t palette[256][4];      /* PALETTE colours; BOX colour in palette[1]           */
  int32_t box[4];               /* BOX: l,t,r,b inclusive (PIL rectangle)              */
  int32_t box_width;
  const b200vit_frame_op* h_ops; /* HOST array [T]; circle stamps of one clip share r   */
  const b200vit_frame_op* d_ops; /* DEVICE array [T] (e.g. written by b200vit_stom_policy); used when h_ops is
                                    NULL.  Not validated by the host: an inconsistent op leaves its frame untouched */
  int32_t d_ops_circle_r;        /* radius of the circle stamps in d_ops (min(h,w)/20, STOM.py:196), or -1        */
} b200vit_overlay;

typedef struct b200vit_frames {
  const uint8_t* d_frames;  /* uint8 [T,H,W,3] (HWC, what PIL / cv2 decode produce)    */
  int32_t t, h, w;          /* T may be odd: last frame repeated (HF videoproc :245-249) */
} b200vit_frames;

/* ------------------------------------------------------------------ hot path */
/* Whole path.  Exactly one of d_pixel_values (bf16 [M, C*tp*p*p], the HF
 * processor's layout) and frames (+ optional overlay) is non-NULL; frames are
 * only valid for a single-grid plan.  d_out: [M/4, out_hidden], bf16
 * (out_f32 = 0) or fp32, in ORIGINAL merged-token order.  d_last_hidden
 * (optional, may be NULL): fp32 [M, D] in window order (HF 5.x
 * last_hidden_state).                                                         */
int b200vit_forward(const b200vit_plan* plan, const b200vit_weights* w, const void* d_pixel_values,
                    const b200vit_frames* frames, const b200vit_overlay* overlay, void* d_out, int out_f32,
                    float* d_last_hidden, void* d_workspace, size_t workspace_bytes, b200vit_stream stream);

/* The forward keeps the fp32 residual stream resident in L2 (a persisting access-policy window on the caller's
 * stream for the duration of the call) when it fits.  That needs the DEVICE-wide limit
 * cudaLimitPersistingL2CacheSize, which the library grows (never shrinks) to the largest residual stream seen and
 * gives back only for a clip whose stream does not fit (long videos run faster with the whole L2 as normal cache).
 * mode 0: never touch the limit (a host application that manages it itself); mode 1 (default): as described.
 * The environment variable B200VIT_L2_PERSIST=0 selects mode 0 at first use.                                    */
int b200vit_set_l2_persist(int mode);

/* Number of kernel launches one b200vit_forward enqueues for this plan.       */
int b200vit_forward_launches(const b200vit_plan* plan, int with_frames);

/* Per-kernel timing of the NEXT forward calls on this plan: a cudaEvent pair is recorded
 * around every launch on the caller's stream.  profile_read waits for the last profiled
 * forward and returns summed milliseconds and launch counts per kernel kind.           */
enum {
  B200VIT_K_OVERLAY_PATCHIFY = 0, B200VIT_K_PATCH_EMBED, B200VIT_K_RMSNORM, B200VIT_K_QKV, B200VIT_K_ATTN_WINDOW,
  B200VIT_K_ATTN_FULL, B200VIT_K_PROJ, B200VIT_K_GATEUP, B200VIT_K_DOWN, B200VIT_K_MERGER_FC1, B200VIT_K_MERGER_FC2,
  B200VIT_K_QKV_WINATTN, /* QKV projection + RoPE + window attention in one kernel (windowed layers, 64-row windows) */
  B200VIT_K_COUNT
};
int b200vit_profile_enable(b200vit_plan* plan, int enable);
int b200vit_profile_read(b200vit_plan* plan, float* h_ms_by_kind, int32_t* h_count_by_kind);

/* ------------------------------------------------------------------ single ops
 * The kernels behind b200vit_forward, exposed for parity tests and profiling. */

/* Overlay only: composited uint8 frames [T,H,W,3] (bit-exact vs PIL).         */
int b200vit_overlay_composite(const b200vit_frames* frames, const b200vit_overlay* overlay, uint8_t* d_out,
                              b200vit_stream stream);
/* Measurement aid: one thread spins for spin_ns (<= 1 ms) and writes {SM cycles, nanoseconds} of that interval to
 * d_out2[2] -- the SM clock under load without an NVML read (bench.py, N > 1).                                     */
int b200vit_clock_probe(uint64_t* d_out2, uint32_t spin_ns, b200vit_stream stream);

/* Frame resize ahead of the overlay: Pillow's bicubic Image.resize on uint8 RGB frames,
 * bit for bit -- what qwen_vl_utils.fetch_image does to every frame after smart_resize (reference call sites
 * app.py:296, :417, utils/dataset.py:76).  d_in [T,h_in,w_in,3] -> d_out [T,h_out,w_out,3]; w_out % 4 == 0.       */
size_t b200vit_resize_workspace_bytes(int32_t t, int32_t h_in, int32_t w_in, int32_t h_out, int32_t w_out);
int b200vit_resize_bicubic(const uint8_t* d_in, int32_t t, int32_t h_in, int32_t w_in, uint8_t* d_out, int32_t h_out,
                           int32_t w_out, void* d_workspace, size_t workspace_bytes, b200vit_stream stream);

/* STOM placement policy on the device (model/STOM.py:72-141 after the tracker): from tracker outputs
 * d_tracks fp32 [T,N,2] (x,y) and d_vis u8 [T,N] (non-zero = visible) to one frame op per frame in d_ops
 * (DEVICE array [T], the d_ops of b200vit_overlay), without a host round trip:
 *   key frame            -> FRAME_LAYER, no shift (:82-86)
 *   mask_shape == 0      -> MAD-filtered mean flow of the visible tracks (:104-131, numpy float32 semantics
 *                           reproduced bit for bit), reduced to the integer shift of STOM.warp (:145-155);
 *                           FRAME_NONE where the reference keeps the frame (:108-110, :122-124, :129-131)
 *   mask_shape != 0      -> STOM.warp_point (:163-203): visible-track mask, closing with the k = min(h,w)/15
 *                           ellipse, centroid, circle of radius min(h,w)/20 in the layer's first non-transparent
 *                           colour with alpha clamped to [96,148]; FRAME_NONE where it returns/raises early
 * d_layer_rgba: the RGBA prompt layer [h,w,4] (only read for mask shapes).  Workspace: b200vit_stom_policy_workspace_bytes.
 * n_points <= 16384.                                                                                        */
size_t b200vit_stom_policy_workspace_bytes(int32_t t_frames, int32_t n_points, int32_t h, int32_t w);
int b200vit_stom_policy(const float* d_tracks, const uint8_t* d_vis, int32_t t_frames, int32_t n_points, int32_t key_idx,
                        int32_t mask_shape, int32_t h, int32_t w, const uint8_t* d_layer_rgba, b200vit_frame_op* d_ops,
                        void* d_workspace, size_t workspace_bytes, b200vit_stream stream);

/* Prompt-layer rasterisers: the mask and scribble prompts of /root/reference/utils/visual_prompt_generator.py
 * (draw_mask :268-274 = ImageDraw.polygon(coords, fill) per contour; draw_scribble :230-252 = one
 * ImageDraw.line([prev, cur], width) per step of a cubic Bezier sampled at 1000 * max(W,H)/anchor points), drawn on
 * the GPU into a uint8 [h, w] palette layer (the d_layer of a B200VIT_LAYER_PALETTE overlay): every pixel Pillow
 * would paint gets `index`, the others are left as they are (zero the layer first).  Coverage is Pillow's bit for
 * bit (integer-truncated vertices, float32 scanline crossings, ROUND_UP / ROUND_DOWN span ends, the corner rules of
 * polygon_generic, ImagingDrawWideLine's quadrilateral; Bresenham for width <= 1).  Vertex lists are HOST arrays of
 * doubles, as ImageDraw hands them to the C library.                                                              */
int b200vit_raster_polygons(const double* h_xy, const int32_t* h_counts, int32_t n_polygons, int32_t h, int32_t w,
                            uint8_t index, uint8_t* d_layer, b200vit_stream stream);
/* n_points - 1 separate two-point lines (x0,y0)-(x1,y1), (x1,y1)-(x2,y2), ... of one width (no joints)           */
int b200vit_raster_lines(const double* h_xy, int32_t n_points, int32_t width, int32_t h, int32_t w, uint8_t index,
                         uint8_t* d_layer, b200vit_stream stream);
/* Host helper: the n_points Bezier samples of draw_scribble (:244-246) for control points h_ctrl8 =
 * {p0x,p0y,p1x,p1y,p2x,p2y,p3x,p3y}, evaluated exactly as the Python expression is (float64, t from np.linspace). */
int b200vit_scribble_points(const double* h_ctrl8, int32_t n_points, double* h_xy_out);

/* Overlay + normalise + patchify: bf16 [M, 3*tp*14*14] in processor order.    */
int b200vit_overlay_patchify(const b200vit_frames* frames, const b200vit_overlay* overlay, int patch, int tps,
                             int merge, void* d_out_bf16, b200vit_stream stream);

enum {
  B200VIT_EPI_STORE_F32 = 0,     /* out_f32[row_map[r]] = acc                                 */
  B200VIT_EPI_QKV_ROPE = 1,      /* B rows head-interleaved [(head, {q,k,v}, 80)] (b200vit_pack_weights); bf16 out [M, 3D] =
                                    Q | K | V head-major, rope(rstd * acc + bias) for Q and K, rstd * acc + bias for V */
  B200VIT_EPI_BIAS_RESIDUAL = 2, /* out_f32 += acc + bias                                     */
  B200VIT_EPI_SWIGLU = 3,        /* bf16 out[:, c/2] = silu(acc[c]+b[c]) * (acc[c+1]+b[c+1])   */
  B200VIT_EPI_BIAS_GELU = 4,     /* bf16 out = gelu_erf(acc + bias)                           */
  B200VIT_EPI_BIAS_BF16 = 5,     /* bf16 out[row_map[r]] = acc + bias                         */
  B200VIT_EPI_BIAS_F32 = 6,      /* f32  out[row_map[r]] = acc + bias                         */
  B200VIT_EPI_BIAS_RESIDUAL_NORM = 7, /* out_f32 += acc + bias, and from the NEW out: d_out_bf16 = bf16(out),
                                    d_rowsq_out partial row sums of out^2 (feeds the next fused RMSNorm) */
  B200VIT_EPI_QKV_ROPE_WINATTN = 8 /* QKV_ROPE followed, inside the epilogue, by the window attention of HF :244-283 for
                                    windows of exactly 64 consecutive rows (M % 64 == 0): bf16 out [M, D] = attention
                                    output, head-major; Q, K, V never leave the SM                                   */
};
/* int32 words of the d_sync scratch (zeroed once; every launch leaves it zeroed) */
#define B200VIT_GEMM_SYNC_INTS 4096
typedef struct b200vit_gemm_args {
  const void* d_a;      /* bf16 [M, K] row-major, lda = K                              */
  const void* d_b;      /* bf16 [N, K] row-major (nn.Linear weight)                    */
  void* d_out;          /* see epilogue                                               */
  const float* d_bias;  /* [N] or NULL                                                */
  const int32_t* d_row_map; /* [M] or NULL (identity)                                  */
  const void* d_rope;   /* QKV_ROPE: fp32 [P, 20, 2]: (cos, sin) of coordinate * inv_freq[k] for coordinates 0..P-1 --
                           HF's own rotary table before the pos_ids gather (modeling :117-130, :382-409)          */
  const int32_t* d_rope_pos; /* QKV_ROPE: int32 [M, 2] (hpos, wpos) of every A row: head dims 0..19 (and 40..59) turn
                           with hpos, 20..39 (and 60..79) with wpos                                               */
  int32_t m, n, k;
  int32_t ldo;          /* leading dimension of out, in elements                       */
  int32_t epilogue;     /* B200VIT_EPI_*                                              */
  /* fused RMSNorm (all optional, NULL/0 = off) */
  void* d_out_bf16;        /* STORE_F32 (optional), BIAS_RESIDUAL_NORM (required): bf16 copy of the fp32 output,
                              same rows (row_map applied) and leading dimension as d_out                      */
  float* d_rowsq_out;      /* same epilogues: [ceil(N/128)][M] per-row sums of out^2 over each 128-column group */
  const float* d_rowsq_in; /* QKV_ROPE, SWIGLU: [rowsq_parts][M] partial sums of squares of the fp32 rows A was
                              cast from; the accumulator row is scaled by rsqrt(sum / K + norm_eps) before the bias */
  int32_t rowsq_parts;
  float norm_eps;
  int32_t* d_sync;         /* BIAS_RESIDUAL_NORM: zeroed int32[B200VIT_GEMM_SYNC_INTS]; enables the balanced
                              (stream-K) decomposition with a fixed, bit-reproducible addition order; NULL = whole tiles */
} b200vit_gemm_args;
/* out = epilogue(A[M,K] * B[N,K]^T): tcgen05/TMEM GEMM fed by TMA.             */
int b200vit_gemm(const b200vit_gemm_args* args, b200vit_stream stream);

/* h_bf16[r,:] = (x[r,:] * rsqrt(mean(x^2)+eps)) * w   (HF modeling :66-71)     */
int b200vit_rmsnorm(const float* d_x, const float* d_w, void* d_out_bf16, int rows, int dim, float eps,
                    b200vit_stream stream);

/* Varlen attention over cu_seqlens segments (HF modeling :244-283): qkv bf16
 * [M, 3*heads*80] (Q|K|V, head-major, RoPE already applied), out bf16 [M, heads*80]. */
int b200vit_attention(const void* d_qkv, void* d_out, const int32_t* h_cu_seqlens, int n_segments, int heads,
                      b200vit_stream stream);

/* fp32 / fp16 / bf16 [n] -> bf16 [n] (the `.type(self.visual.dtype)` at HF :1148) */
int b200vit_cast_to_bf16(const void* d_in, int in_dtype /*0 f32, 1 f16, 2 bf16*/, void* d_out, int64_t n,
                         b200vit_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* B200VIT_H_ */
