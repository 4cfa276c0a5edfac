"""Drop-in replacement for HF ``Qwen2_5_VisionTransformerPretrainedModel`` backed by
libb200vit.so.

Same constructor config, same ``state_dict`` keys/shapes, same call signature
``forward(hidden_states [M,1176], grid_thw [n,3]) -> merged embeddings`` as the
module the reference calls at /root/reference/model/qwen_2_5_vl_sam2.py:182-200
(through HF ``get_video_features``, modeling_qwen2_5_vl.py:1137-1177), plus the
fused entry ``forward_frames`` that starts from uint8 frames and a STOM overlay.

PyTorch is used for device memory, streams and (optional) CUDA-graph capture only;
all arithmetic runs in the hand-written sm_100a kernels.  There is no fallback.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import nn

from . import _lib
from .overlay import OverlaySpec


def _norm_device(device) -> torch.device:
    """torch.device with an explicit CUDA index ("cuda" -> the current device), so device comparisons are exact."""
    d = torch.device(device)
    if d.type == "cuda" and d.index is None:
        d = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
    return d


class _Holder(nn.Module):
    """Parameter holder with the HF names (weight / bias)."""

    def __init__(self, w_shape, bias: bool, dtype, device, ones: bool = False):
        super().__init__()
        init = torch.ones if ones else torch.zeros
        self.weight = nn.Parameter(init(w_shape, dtype=dtype, device=device), requires_grad=False)
        if bias:
            self.bias = nn.Parameter(torch.zeros(w_shape[0], dtype=dtype, device=device), requires_grad=False)


class _Plan:
    """Owns one b200vit_plan (host + device index tables).  Immutable once uploaded, so ONE plan per grid serves every
    stream / slot.  The activation workspace is NOT per plan: the tower keeps one per slot, sized for the largest grid
    seen, so a service that sees many resolutions does not accumulate gigabytes of idle workspaces."""

    def __init__(self, grid: Tuple[Tuple[int, int, int], ...], cfg_c, device):
        self.grid = grid
        arr = (C.c_int64 * (3 * len(grid)))(*[v for g in grid for v in g])
        handle = C.c_void_p()
        _lib.check(_lib.lib().b200vit_plan_create(arr, len(grid), C.byref(cfg_c), C.byref(handle)), "plan_create")
        self.handle = handle
        self.m = int(sum(t * h * w for t, h, w in grid))
        self.ws_bytes = int(_lib.lib().b200vit_workspace_bytes(handle))
        self.device = device
        self.graphs: Dict[str, tuple] = {}

    def get(self, which: int, dtype) -> np.ndarray:
        n = int(_lib.lib().b200vit_plan_get(self.handle, which, None, 0))
        out = np.empty(n // np.dtype(dtype).itemsize, dtype=dtype)
        _lib.lib().b200vit_plan_get(self.handle, which, out.ctypes.data_as(C.c_void_p), n)
        return out

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().b200vit_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class _Output:
    """Minimal stand-in for transformers' BaseModelOutputWithPooling (5.x return style)."""

    def __init__(self, pooler_output, last_hidden_state=None):
        self.pooler_output = pooler_output
        self.last_hidden_state = last_hidden_state

    def __getitem__(self, i):
        return (self.last_hidden_state, self.pooler_output)[i]


class B200VisionTower(nn.Module):
    def __init__(self, config, device="cuda", dtype=torch.bfloat16, return_dict: Optional[bool] = None,
                 use_cuda_graph: bool = False, output_fp32: bool = False, max_plans: int = 32):
        super().__init__()
        self.max_plans = max(int(max_plans), 1)
        g = (lambda k, d=None: getattr(config, k, d)) if not isinstance(config, dict) else (lambda k, d=None: config.get(k, d))
        self.config = config
        self.depth = g("depth", 32)
        self.hidden_size = g("hidden_size", 1280)
        self.intermediate_size = g("intermediate_size", 3420)
        self.num_heads = g("num_heads", 16)
        self.out_hidden_size = g("out_hidden_size", 3584)
        self.patch_size = g("patch_size", 14)
        self.temporal_patch_size = g("temporal_patch_size", 2)
        self.spatial_merge_size = g("spatial_merge_size", 2)
        self.spatial_merge_unit = self.spatial_merge_size ** 2
        self.window_size = g("window_size", 112)
        self.in_channels = g("in_channels", 3)
        self.fullatt_block_indexes = list(g("fullatt_block_indexes", [7, 15, 23, 31]))
        if g("hidden_act", "silu") != "silu":
            raise ValueError("only the SiLU-gated MLP of Qwen2.5-VL is implemented")
        self._dtype = dtype
        self._device = _norm_device(device)
        self.use_cuda_graph = use_cuda_graph
        self.output_fp32 = output_fp32
        if return_dict is None:  # transformers >= 5 returns an object with .pooler_output, 4.49 a tensor
            try:
                import transformers
                return_dict = int(transformers.__version__.split(".")[0]) >= 5
            except Exception:
                return_dict = False
        self.return_dict = return_dict

        d, i, o, u = self.hidden_size, self.intermediate_size, self.out_hidden_size, self.spatial_merge_unit
        mk = lambda shape, bias, ones=False: _Holder(shape, bias, dtype, self._device, ones)
        self.patch_embed = nn.Module()
        self.patch_embed.proj = mk((d, self.in_channels, self.temporal_patch_size, self.patch_size, self.patch_size), False)
        self.blocks = nn.ModuleList()
        for _ in range(self.depth):
            b = nn.Module()
            b.norm1, b.norm2 = mk((d,), False, True), mk((d,), False, True)
            b.attn = nn.Module()
            b.attn.qkv, b.attn.proj = mk((3 * d, d), True), mk((d, d), True)
            b.mlp = nn.Module()
            b.mlp.gate_proj, b.mlp.up_proj, b.mlp.down_proj = mk((i, d), True), mk((i, d), True), mk((d, i), True)
            self.blocks.append(b)
        self.merger = nn.Module()
        self.merger.ln_q = mk((d,), False, True)
        self.merger.mlp = nn.ModuleDict({"0": mk((u * d, u * d), True), "2": mk((o, u * d), True)})

        self._cfg_c = _lib.Cfg()
        for k, v in dict(depth=self.depth, hidden=d, intermediate=i, heads=self.num_heads, out_hidden=o,
                         patch=self.patch_size, temporal_patch=self.temporal_patch_size, merge=self.spatial_merge_size,
                         window=self.window_size, in_channels=self.in_channels,
                         n_fullatt=len(self.fullatt_block_indexes)).items():
            setattr(self._cfg_c, k, int(v))
        for j, v in enumerate(self.fullatt_block_indexes):
            self._cfg_c.fullatt[j] = int(v)
        self._packed = None
        self._plans: "OrderedDict[tuple, _Plan]" = OrderedDict()   # LRU, at most max_plans entries
        self._workspaces: Dict[int, tuple] = {}                     # slot -> (tensor, aligned pointer, usable bytes)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    # ---- HF-style attributes
    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._device

    def _invalidate(self):
        self._packed = None
        for q in self._plans.values():      # captured graphs replay the old packed buffer
            q.graphs.clear()

    def invalidate(self):
        """Call after editing a parameter in place (``tower.blocks[0].attn.qkv.bias.add_(...)``): the packed device
        copies are rebuilt by the next forward.  ``load_state_dict`` / ``.to()`` / ``.half()`` do this themselves."""
        self._invalidate()

    def _apply(self, fn, *args, **kwargs):
        """`.to()`, `.cuda()`, `.half()` ...: the HF-named parameters move or change dtype, so the packed copies (and,
        on a device change, every plan and workspace) are rebuilt lazily by the next forward; `dtype` / `device`
        follow the parameters, as they do on the HF module."""
        out = super()._apply(fn, *args, **kwargs)
        p = next(self.parameters(), None)
        if p is not None and _norm_device(p.device) != self._device:
            self._device = _norm_device(p.device)
            self._plans.clear()
            self._workspaces.clear()
        if p is not None and p.dtype.is_floating_point and p.dtype != self._dtype:
            self._dtype = p.dtype
        self._invalidate()
        return out

    @classmethod
    def from_hf(cls, hf_tower, **kw):
        """Build from an instantiated HF tower (copies its weights)."""
        dev = kw.pop("device", "cuda")
        m = cls(hf_tower.config, device=dev, **kw)
        sd = {k: v for k, v in hf_tower.state_dict().items() if "inv_freq" not in k}
        m.load_state_dict(sd)
        return m

    # ---- weight packing (once): b200vit_pack_weights does the bf16 casts, the RMSNorm gamma fold, the gate/up
    # interleave and the I -> Ipad padding on the device; this method only hands it the HF-named tensors.
    @torch.no_grad()
    def pack_weights(self):
        dev = self._device
        params = [p for p in self.parameters()]
        dt = params[0].dtype
        code = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}.get(dt)
        keep = []

        def ptr(t):
            if code is None or t.dtype != dt or t.device != dev or not t.is_contiguous():
                t = t.detach().to(device=dev, dtype=dt if code is not None else torch.float32).contiguous()
                keep.append(t)
            return t.data_ptr()

        raw_layers = (_lib.RawLayer * self.depth)()
        for li, b in enumerate(self.blocks):
            r = raw_layers[li]
            r.norm1_w, r.norm2_w = ptr(b.norm1.weight), ptr(b.norm2.weight)
            r.qkv_w, r.qkv_b = ptr(b.attn.qkv.weight), ptr(b.attn.qkv.bias)
            r.proj_w, r.proj_b = ptr(b.attn.proj.weight), ptr(b.attn.proj.bias)
            r.gate_w, r.gate_b = ptr(b.mlp.gate_proj.weight), ptr(b.mlp.gate_proj.bias)
            r.up_w, r.up_b = ptr(b.mlp.up_proj.weight), ptr(b.mlp.up_proj.bias)
            r.down_w, r.down_b = ptr(b.mlp.down_proj.weight), ptr(b.mlp.down_proj.bias)
        raw = _lib.RawWeights()
        raw.dtype = code if code is not None else 0
        raw.patch_w = ptr(self.patch_embed.proj.weight)
        raw.layers = raw_layers
        raw.merger_ln_w = ptr(self.merger.ln_q.weight)
        raw.merger_fc1_w, raw.merger_fc1_b = ptr(self.merger.mlp["0"].weight), ptr(self.merger.mlp["0"].bias)
        raw.merger_fc2_w, raw.merger_fc2_b = ptr(self.merger.mlp["2"].weight), ptr(self.merger.mlp["2"].bias)
        nbytes = int(_lib.lib().b200vit_packed_weights_bytes(C.byref(self._cfg_c)))
        with torch.cuda.device(dev):
            buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            base = buf.data_ptr() + ((-buf.data_ptr()) % 256)
            w, layers = _lib.Weights(), (_lib.LayerWeights * self.depth)()
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(_lib.lib().b200vit_pack_weights(C.byref(self._cfg_c), C.byref(raw), base, nbytes, C.byref(w), layers,
                                                       stream), "pack_weights")
        self._packed = (w, layers, buf)
        return w

    def _weights(self):
        if self._packed is None:
            self.pack_weights()
        return self._packed[0]

    # ---- plans
    def plan_for(self, grid_thw) -> _Plan:
        """Plan (index tables, workspace layout) for this grid; shared by every stream / slot."""
        if isinstance(grid_thw, torch.Tensor):
            grid = tuple(tuple(int(v) for v in row) for row in grid_thw.detach().cpu().tolist())
        else:
            grid = tuple(tuple(int(v) for v in row) for row in np.asarray(grid_thw).reshape(-1, 3).tolist())
        p = self._plans.get(grid)
        if p is None:
            p = _Plan(grid, self._cfg_c, self._device)
            self._plans[grid] = p
            while len(self._plans) > self.max_plans:       # least recently used first; never the one just made
                self._plans.popitem(last=False)
        else:
            self._plans.move_to_end(grid)
        return p

    def _workspace(self, plan: _Plan, slot: int = 0) -> int:
        """1024-byte aligned workspace pointer of `slot`, grown (never shrunk) to the largest request.  Clips that are
        in flight at the same time (different CUDA streams) must use different slots."""
        ws = self._workspaces.get(slot)
        if ws is None or ws[2] < plan.ws_bytes:
            t = torch.empty(plan.ws_bytes + 1024, dtype=torch.uint8, device=self._device)
            ptr = t.data_ptr() + ((-t.data_ptr()) % 1024)
            self._workspaces[slot] = ws = (t, ptr, plan.ws_bytes)
            for q in self._plans.values():                  # captured graphs point into the old buffer
                q.graphs.clear()
        return ws[1]

    def _check_device(self, t: torch.Tensor, what: str):
        if not t.is_cuda or _norm_device(t.device) != self._device:
            raise ValueError(f"B200VisionTower: {what} must be a CUDA tensor on {self._device} (got {t.device}); "
                             "there is no CPU path and no implicit cross-device copy")

    # ---- the hot path
    def _run(self, plan: _Plan, pixel_values, frames_c, overlay_c, out, last_hidden, slot: int = 0):
        with torch.cuda.device(self._device):      # allocations, lazy uploads and launches go to the tower's device
            stream = torch.cuda.current_stream(self._device).cuda_stream
            rc = _lib.lib().b200vit_forward(
                plan.handle, C.byref(self._weights()),
                pixel_values.data_ptr() if pixel_values is not None else None,
                C.byref(frames_c) if frames_c is not None else None,
                C.byref(overlay_c) if overlay_c is not None else None,
                out.data_ptr(), 1 if out.dtype == torch.float32 else 0,
                last_hidden.data_ptr() if last_hidden is not None else None,
                self._workspace(plan, slot), plan.ws_bytes, stream)
        _lib.check(rc, "b200vit_forward")

    def _wrap(self, out, last_hidden):
        if self.return_dict:
            return _Output(out, last_hidden)
        return out

    def _check_out(self, out, shape):
        if tuple(out.shape) != shape or not out.is_contiguous() or out.dtype not in (torch.bfloat16, torch.float32):
            raise ValueError(f"out must be a contiguous bf16/fp32 tensor of shape {shape}")
        self._check_device(out, "out")

    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor, grid_thw, output_last_hidden_state: bool = False,
                out: Optional[torch.Tensor] = None, **kwargs):
        """hidden_states: [M, C*tp*p*p] float (bf16/fp16/fp32), CUDA; grid_thw: [n,3].
        `out` (optional): contiguous [M/4, out_hidden] bf16/fp32 destination, e.g. the placeholder rows of the LLM's
        `inputs_embeds` (see `splice_span`): the merger epilogue then writes the visual tokens in place and HF's
        boolean-mask `masked_scatter` (modeling_qwen2_5_vl.py:1309-1315) becomes unnecessary."""
        self._check_device(hidden_states, "hidden_states")
        plan = self.plan_for(grid_thw)
        kpe = self.in_channels * self.temporal_patch_size * self.patch_size ** 2
        if hidden_states.dim() != 2 or hidden_states.shape[0] != plan.m or hidden_states.shape[1] != kpe:
            raise ValueError(f"hidden_states must be [{plan.m}, {kpe}] for grid_thw {plan.grid}, got {tuple(hidden_states.shape)}")
        x = hidden_states.contiguous()
        if x.dtype != torch.bfloat16:
            code = {torch.float32: 0, torch.float16: 1}.get(x.dtype)
            if code is None:
                raise ValueError(f"unsupported pixel dtype {x.dtype}")
            with torch.cuda.device(self._device):
                xb = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
                stream = torch.cuda.current_stream(self._device).cuda_stream
                _lib.check(_lib.lib().b200vit_cast_to_bf16(x.data_ptr(), code, xb.data_ptr(), x.numel(), stream), "cast")
            x = xb
        out_dtype = torch.float32 if self.output_fp32 else self._dtype
        if out_dtype not in (torch.bfloat16, torch.float32):
            out_dtype = torch.bfloat16                      # a .half() tower still computes and returns bf16
        if self.use_cuda_graph and not output_last_hidden_state and out is None:
            return self._wrap(self._graph_forward(plan, x, out_dtype), None)
        shape = (plan.m // self.spatial_merge_unit, self.out_hidden_size)
        if out is None:
            out = torch.empty(shape, dtype=out_dtype, device=x.device)
        else:
            self._check_out(out, shape)
        last = torch.empty(plan.m, self.hidden_size, dtype=torch.float32, device=x.device) if output_last_hidden_state else None
        self._run(plan, x, None, None, out, last)
        return self._wrap(out, last)

    def _graph_forward(self, plan: _Plan, x, out_dtype):
        key = f"pv-{out_dtype}"
        g = plan.graphs.get(key)
        if g is None:
            static_in = torch.empty_like(x)
            static_out = torch.empty(plan.m // self.spatial_merge_unit, self.out_hidden_size, dtype=out_dtype, device=x.device)
            static_in.copy_(x)
            s = torch.cuda.Stream(self._device)
            s.wait_stream(torch.cuda.current_stream(self._device))
            with torch.cuda.stream(s):
                for _ in range(2):  # warm-up outside capture: lazy uploads, func attributes
                    self._run(plan, static_in, None, None, static_out, None)
            torch.cuda.current_stream(self._device).wait_stream(s)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._run(plan, static_in, None, None, static_out, None)
            g = (graph, static_in, static_out)
            plan.graphs[key] = g
        graph, static_in, static_out = g
        static_in.copy_(x)
        graph.replay()
        return static_out.clone()   # the graph's own buffer is overwritten by the next replay (two clips of one shape)

    @torch.no_grad()
    def forward_frames(self, frames_u8: torch.Tensor, overlay: Optional[OverlaySpec] = None, grid_thw=None,
                       out: Optional[torch.Tensor] = None, slot: int = 0):
        """Fused entry: uint8 frames [T,H,W,3] (CUDA) + optional STOM overlay -> merged embeddings.
        Equivalent to PIL overlay -> Qwen2VLVideoProcessor(do_resize=False) -> tower.  `slot` selects the workspace:
        clips in flight on different CUDA streams use different slots."""
        if not frames_u8.is_cuda or frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[3] != 3:
            raise ValueError("frames must be a CUDA uint8 tensor [T,H,W,3]")
        self._check_device(frames_u8, "frames")
        t, h, w, _ = frames_u8.shape
        if grid_thw is None:
            tps = self.temporal_patch_size
            grid_thw = [[(t + tps - 1) // tps, h // self.patch_size, w // self.patch_size]]
        plan = self.plan_for(grid_thw)
        fr = frames_u8.contiguous()
        fc = _lib.Frames(fr.data_ptr(), t, h, w)
        if overlay is not None:
            overlay.validate(h, w, self._device)   # Image.alpha_composite raises on a size mismatch; so do we
        oc = overlay.to_c(t) if overlay is not None else None
        out_dtype = torch.float32 if self.output_fp32 else self._dtype
        if out_dtype not in (torch.bfloat16, torch.float32):
            out_dtype = torch.bfloat16
        shape = (plan.m // self.spatial_merge_unit, self.out_hidden_size)
        if out is None:
            out = torch.empty(shape, dtype=out_dtype, device=fr.device)
        else:
            self._check_out(out, shape)
        self._run(plan, None, fc, oc, out, None, slot)
        return self._wrap(out, None)

    def profile(self, grid_thw, enable: bool):
        """Per-kernel cudaEvent timing of the following forwards on this grid's plan."""
        _lib.check(_lib.lib().b200vit_profile_enable(self.plan_for(grid_thw).handle, 1 if enable else 0), "profile_enable")

    def profile_read(self, grid_thw):
        """{kind: (total_ms, launches)} for the last profiled forward."""
        n = len(_lib.KERNEL_KINDS)
        ms, cnt = (C.c_float * n)(), (C.c_int32 * n)()
        _lib.check(_lib.lib().b200vit_profile_read(self.plan_for(grid_thw).handle, ms, cnt), "profile_read")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(_lib.KERNEL_KINDS)}

    def launches_per_forward(self, grid_thw, with_frames=False) -> int:
        return int(_lib.lib().b200vit_forward_launches(self.plan_for(grid_thw).handle, 1 if with_frames else 0))


def splice_span(input_ids: torch.Tensor, token_id: int):
    """(start, length) of the run of visual placeholder tokens in a [1, L] / [L] `input_ids`.
    Qwen2.5-VL emits all placeholders of one video as one contiguous run (vision_start, N x video_token, vision_end), so
    `tower(pixel_values, grid, out=inputs_embeds[0, start:start + length])` replaces get_video_features + masked_scatter.
    Raises if the placeholders are not contiguous (several videos: call once per video with its own span)."""
    ids = input_ids.reshape(-1)
    pos = torch.nonzero(ids == token_id).reshape(-1)
    if pos.numel() == 0:
        raise ValueError("no visual placeholder tokens in input_ids")
    start, length = int(pos[0]), int(pos.numel())
    if int(pos[-1]) - start + 1 != length:
        raise ValueError("visual placeholder tokens are not one contiguous run")
    return start, length


def install(model, **kw):
    """Swap the HF vision tower inside a Qwen2.5-VL / UniGR model for the B200 one.
    Handles both attribute layouts: ``model.model.visual`` (transformers 5.x) and
    ``model.visual`` (4.49, used by /root/reference/train_joint.py:190)."""
    holder = model.model if hasattr(getattr(model, "model", None), "visual") else model
    hf_tower = holder.visual
    p0 = next(hf_tower.parameters())
    kw.setdefault("dtype", p0.dtype)                        # `tower.dtype` is what HF casts pixel_values to (:1148)
    tower = B200VisionTower.from_hf(hf_tower, device=p0.device, **kw)
    holder.visual = tower
    return tower
