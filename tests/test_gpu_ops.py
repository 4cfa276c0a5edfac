"""Per-kernel parity on the B200: every CUDA kernel, called through the C ABI,
against the CPU oracle / an fp32 restatement on the same seeded inputs."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import rga3_release_b200 as vit
from rga3_release_b200 import _lib
from oracle import overlay_ref, patchify_ref, tower_ref

DEV = "cuda"


def _stream():
    return torch.cuda.current_stream().cuda_stream


def make_rope(m, seed, side=50):
    """Random rotary inputs in the form the QKV epilogue reads them (HF :382-409 before / after the pos_ids gather):
    a [side, 20, 2] fp32 (cos, sin) table by coordinate and [m, 2] int32 (hpos, wpos) per row.  Also returns the
    gathered [m, 40] cos / sin an HF-style reference needs."""
    g = torch.Generator().manual_seed(seed)
    ang = torch.randn(side, 20, generator=g) * 3.0
    pos = torch.randint(0, side, (m, 2), generator=g, dtype=torch.int32)
    table = torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()
    full = torch.cat([ang[pos[:, 0].long()], ang[pos[:, 1].long()]], dim=-1)      # [m, 40]
    return (table.to(DEV), pos.to(DEV)), full.cos(), full.sin()


def run_gemm(a, b, epi, out, bias=None, row_map=None, rope=None, ldo=None, out_bf16=None, rowsq_out=None,
             rowsq_in=None, eps=1e-6, sync=None):
    g = _lib.GemmArgs()
    g.d_a, g.d_b, g.d_out = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.d_bias = bias.data_ptr() if bias is not None else None
    g.d_row_map = row_map.data_ptr() if row_map is not None else None
    if rope is not None:
        g.d_rope, g.d_rope_pos = rope[0].data_ptr(), rope[1].data_ptr()
    g.m, g.k = a.shape
    g.n = b.shape[0]
    g.ldo = ldo if ldo is not None else out.shape[1]
    g.epilogue = epi
    g.d_out_bf16 = out_bf16.data_ptr() if out_bf16 is not None else None
    g.d_rowsq_out = rowsq_out.data_ptr() if rowsq_out is not None else None
    if rowsq_in is not None:
        g.d_rowsq_in, g.rowsq_parts, g.norm_eps = rowsq_in.data_ptr(), rowsq_in.shape[0], eps
    g.d_sync = sync.data_ptr() if sync is not None else None
    _lib.check(_lib.lib().b200vit_gemm(C.byref(g), _stream()), "gemm")
    torch.cuda.synchronize()


def head_major(w, d):
    """HF qkv rows [q heads | k heads | v heads] -> the kernels' head-major order [(head, {q,k,v}, 80)]
    (what b200vit_pack_weights does to attn.qkv.weight / bias)."""
    nh = d // 80
    return w.reshape(3, nh, 80, *w.shape[1:]).transpose(0, 1).reshape(w.shape).contiguous()


def rowsq_parts(x):
    """[ceil(N/128), M] per-row sums of squares over each 128-column group (fp64 reference)."""
    m, n = x.shape
    parts = (n + 127) // 128
    out = torch.zeros(parts, m, dtype=torch.float64)
    xd = x.double().cpu()
    for i in range(parts):
        out[i] = (xd[:, i * 128:(i + 1) * 128] ** 2).sum(-1)
    return out


def rnd(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(DEV)


def close(out, ref, tol=2e-2):
    out, ref = out.float().cpu(), ref.float().cpu()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-12
    assert err <= tol * scale, f"max err {err} vs scale {scale}"


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (300, 160, 1176), (1000, 1280, 160), (4096, 1280, 1280), (77, 512, 3456)])
def test_gemm_store_f32_rowmap(m, n, k):
    a, b = rnd((m, k), 1), rnd((n, k), 2, 0.05)
    perm = torch.randperm(m, generator=torch.Generator().manual_seed(3)).to(torch.int32).to(DEV)
    out = torch.zeros(m, n, dtype=torch.float32, device=DEV)
    run_gemm(a, b, _lib.EPI_STORE_F32, out, row_map=perm)
    ref = torch.empty_like(out)
    ref[perm.long()] = a.float() @ b.float().t()
    close(out, ref, 1e-3)


def test_gemm_exact_small_integers():
    # integer-valued bf16 inputs: fp32 accumulation is exact, so the result must be bit-exact
    g = torch.Generator().manual_seed(0)
    a = torch.randint(-4, 5, (256, 1280), generator=g).to(torch.bfloat16).to(DEV)
    b = torch.randint(-4, 5, (512, 1280), generator=g).to(torch.bfloat16).to(DEV)
    out = torch.zeros(256, 512, dtype=torch.float32, device=DEV)
    run_gemm(a, b, _lib.EPI_STORE_F32, out)
    assert torch.equal(out, a.float() @ b.float().t())


@pytest.mark.parametrize("m,d", [(200, 160), (1024, 1280)])
def test_gemm_qkv_rope(m, d):
    nh = d // 80
    a, b = rnd((m, d), 4), rnd((3 * d, d), 5, 0.05)
    bias = rnd((3 * d,), 6, 0.1, torch.float32)
    rope, cos, sin = make_rope(m, 7)
    out = torch.zeros(m, 3 * d, dtype=torch.bfloat16, device=DEV)
    run_gemm(a, head_major(b, d), _lib.EPI_QKV_ROPE, out, bias=head_major(bias, d), rope=rope)
    qkv = (a.float() @ b.float().t() + bias).cpu().reshape(m, 3, nh, 80)
    c80 = torch.cat([cos, cos], -1).cpu()
    s80 = torch.cat([sin, sin], -1).cpu()
    q, k = tower_ref.rope_ref(qkv[:, 0], qkv[:, 1], c80, s80)
    ref = torch.stack([q, k, qkv[:, 2]], 1).reshape(m, 3 * d)
    close(out, ref, 1e-2)


@pytest.mark.parametrize("m,n,k", [(300, 160, 256), (2048, 1280, 3456)])
def test_gemm_bias_residual(m, n, k):
    a, b = rnd((m, k), 8), rnd((n, k), 9, 0.05)
    bias = rnd((n,), 10, 0.1, torch.float32)
    x0 = rnd((m, n), 11, 1.0, torch.float32)
    x = x0.clone()
    run_gemm(a, b, _lib.EPI_BIAS_RESIDUAL, x, bias=bias)
    close(x, x0 + a.float() @ b.float().t() + bias, 1e-3)


@pytest.mark.parametrize("m,n,k", [(300, 160, 1176), (4096, 1280, 1176)])
def test_gemm_store_f32_emits_norm_inputs(m, n, k):
    """Patch-embed epilogue of the fused-RMSNorm path: besides the fp32 rows (window order) it writes their bf16 copy
    and the per-128-column partial row sums of squares."""
    a, b = rnd((m, k), 1), rnd((n, k), 2, 0.05)
    perm = torch.randperm(m, generator=torch.Generator().manual_seed(3)).to(torch.int32).to(DEV)
    out = torch.zeros(m, n, dtype=torch.float32, device=DEV)
    ob = torch.zeros(m, n, dtype=torch.bfloat16, device=DEV)
    parts = (n + 127) // 128
    rs = torch.zeros(parts, m, dtype=torch.float32, device=DEV)
    run_gemm(a, b, _lib.EPI_STORE_F32, out, row_map=perm, out_bf16=ob, rowsq_out=rs)
    ref = torch.empty_like(out)
    ref[perm.long()] = a.float() @ b.float().t()
    close(out, ref, 1e-3)
    assert torch.equal(ob, out.to(torch.bfloat16))                      # the copy is the rounding of what was stored
    want = rowsq_parts(out)
    assert torch.allclose(rs.double().cpu(), want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("m,n,k,stream_k", [(300, 160, 256, False), (2048, 1280, 3456, False), (8192, 1280, 1280, True),
                                            (8192, 1280, 3456, True), (5000, 1280, 1280, True)])
def test_gemm_bias_residual_norm(m, n, k, stream_k):
    """x += A W^T + b with the new x in hand: fp32 x, its bf16 copy and the row-square partials; with the stream-K
    hand-over scratch the result is bit-identical run to run and to the whole-tile schedule up to fp32 add order."""
    a, b = rnd((m, k), 8), rnd((n, k), 9, 0.05)
    bias = rnd((n,), 10, 0.1, torch.float32)
    x0 = rnd((m, n), 11, 1.0, torch.float32)
    sync = torch.zeros(_lib.GEMM_SYNC_INTS, dtype=torch.int32, device=DEV) if stream_k else None
    parts = (n + 127) // 128
    runs = []
    for _ in range(3 if stream_k else 1):
        x = x0.clone()
        xb = torch.zeros(m, n, dtype=torch.bfloat16, device=DEV)
        rs = torch.full((parts, m), -1.0, dtype=torch.float32, device=DEV)
        run_gemm(a, b, _lib.EPI_BIAS_RESIDUAL_NORM, x, bias=bias, out_bf16=xb, rowsq_out=rs, sync=sync)
        runs.append((x, xb, rs))
        if sync is not None:
            assert int(sync.abs().sum()) == 0                            # every launch leaves the flags zeroed
    x, xb, rs = runs[0]
    close(x, x0 + a.float() @ b.float().t() + bias, 1e-3)
    assert torch.equal(xb, x.to(torch.bfloat16))
    assert torch.allclose(rs.double().cpu(), rowsq_parts(x), rtol=1e-5, atol=1e-6)
    for x2, xb2, rs2 in runs[1:]:
        assert torch.equal(x2, x) and torch.equal(xb2, xb) and torch.equal(rs2, rs)


def test_gemm_fused_rmsnorm_consumers():
    """QKV and SwiGLU epilogues with d_rowsq_in: equal to RMSNorm(x, gamma) @ W^T computed the HF way, when gamma is
    folded into W's columns and the A operand is the bf16 copy of the raw x (HF modeling :57-71)."""
    m, d, ipad = 1000, 1280, 256
    x = rnd((m, d), 20, 3.0, torch.float32) * (1 + torch.arange(m, device=DEV).float().unsqueeze(1) / 100)  # rows of very different scale
    gamma = 1 + 0.1 * rnd((d,), 21, 1.0, torch.float32)
    xb = x.to(torch.bfloat16)
    rs_in = rowsq_parts(x).float().to(DEV)
    normed = (x.double() * torch.rsqrt((x.double() ** 2).mean(-1, keepdim=True) + 1e-6)) * gamma.double()
    # qkv
    w = rnd((3 * d, d), 22, 0.05, torch.float32)
    bias = rnd((3 * d,), 23, 0.1, torch.float32)
    rope, cos, sin = make_rope(m, 24)
    out = torch.zeros(m, 3 * d, dtype=torch.bfloat16, device=DEV)
    run_gemm(xb, head_major((w * gamma).to(torch.bfloat16), d), _lib.EPI_QKV_ROPE, out, bias=head_major(bias, d), rope=rope,
             rowsq_in=rs_in)
    qkv = (normed @ w.double().t() + bias.double()).float().cpu().reshape(m, 3, d // 80, 80)
    c80, s80 = torch.cat([cos, cos], -1).cpu(), torch.cat([sin, sin], -1).cpu()
    q, k = tower_ref.rope_ref(qkv[:, 0], qkv[:, 1], c80, s80)
    close(out, torch.stack([q, k, qkv[:, 2]], 1).reshape(m, 3 * d), 1e-2)
    # gate/up
    w2 = rnd((2 * ipad, d), 25, 0.05, torch.float32)
    b2 = rnd((2 * ipad,), 26, 0.1, torch.float32)
    out2 = torch.zeros(m, ipad, dtype=torch.bfloat16, device=DEV)
    run_gemm(xb, (w2 * gamma).to(torch.bfloat16), _lib.EPI_SWIGLU, out2, bias=b2, ldo=ipad, rowsq_in=rs_in)
    z = (normed @ w2.double().t() + b2.double()).float()
    close(out2, torch.nn.functional.silu(z[:, 0::2]) * z[:, 1::2], 1e-2)


@pytest.mark.parametrize("m,d", [(192, 160), (2048, 1280), (8192 + 64, 320)])
def test_gemm_qkv_rope_window_attention_fused(m, d):
    """QKV_ROPE_WINATTN: projection + RoPE + softmax(Q K^T / sqrt(80)) V over windows of 64 consecutive rows, all in
    the GEMM epilogue, against the unfused restatement (HF :231-261 with cu_window_seqlens = 0, 64, 128, ...)."""
    nh = d // 80
    x = rnd((m, d), 30, 2.0, torch.float32)
    gamma = 1 + 0.1 * rnd((d,), 31, 1.0, torch.float32)
    w, bias = rnd((3 * d, d), 32, 0.03, torch.float32), rnd((3 * d,), 33, 0.1, torch.float32)   # logits of a few units
    rope, cos, sin = make_rope(m, 34)
    xb = x.to(torch.bfloat16)
    rs_in = rowsq_parts(x).float().to(DEV)
    out = torch.zeros(m, d, dtype=torch.bfloat16, device=DEV)
    run_gemm(xb, head_major((w * gamma).to(torch.bfloat16), d), _lib.EPI_QKV_ROPE_WINATTN, out, bias=head_major(bias, d),
             rope=rope, rowsq_in=rs_in, ldo=d)
    normed = (x.double() * torch.rsqrt((x.double() ** 2).mean(-1, keepdim=True) + 1e-6)) * gamma.double()
    qkv = (normed @ w.double().t() + bias.double()).float().cpu().reshape(m, 3, nh, 80)
    c80, s80 = torch.cat([cos, cos], -1), torch.cat([sin, sin], -1)
    q, k = tower_ref.rope_ref(qkv[:, 0], qkv[:, 1], c80, s80)
    v = qkv[:, 2]
    ref = torch.empty(m, nh, 80)
    for w0 in range(0, m, 64):
        qq, kk, vv = (t[w0:w0 + 64].transpose(0, 1) for t in (q, k, v))          # [nh, 64, 80]
        p = torch.softmax(qq @ kk.transpose(1, 2) / 80 ** 0.5, dim=-1)
        ref[w0:w0 + 64] = (p @ vv).transpose(0, 1)
    close(out, ref.reshape(m, d), 1e-2)
    with pytest.raises(ValueError):                       # windows must tile the rows exactly
        run_gemm(xb[:100], head_major((w * gamma).to(torch.bfloat16), d), _lib.EPI_QKV_ROPE_WINATTN, out[:100],
                 bias=head_major(bias, d), rope=rope, ldo=d)


def test_pack_weights_c_abi_matches_torch_restatement():
    """b200vit_pack_weights (host or device pointers, fp32/bf16 sources) == the packing rule restated in torch."""
    from oracle import hf_ref
    cfgk = dict(hf_ref.CFG_SMALL)
    cfg = tower_ref.TowerCfg(**cfgk)
    sd = hf_ref.make_state_dict(cfg, seed=4)
    d, i = cfg.hidden_size, cfg.intermediate_size
    ipad = (i + 127) // 128 * 128
    for dt, on_dev in ((torch.float32, False), (torch.bfloat16, True)):
        t = vit.B200VisionTower(cfgk, device=DEV if on_dev else "cpu", dtype=dt, return_dict=False)
        t.load_state_dict(sd)
        t._device = torch.device(DEV, torch.cuda.current_device())      # host-resident parameters, packed buffer on the GPU
        w = t.pack_weights()
        torch.cuda.synchronize()
        lw = w.layers[1]
        src = {k: v.to(dt).float() for k, v in sd.items()}

        def dev_tensor(ptr, shape, dtype):
            n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            view = type("DevView", (), {"__cuda_array_interface__": dict(shape=(n,), typestr="|u1", data=(int(ptr), False), version=2)})()
            return torch.as_tensor(view, device=DEV).clone().view(dtype).reshape(shape)

        g1, g2 = src["blocks.1.norm1.weight"], src["blocks.1.norm2.weight"]
        qkv = dev_tensor(lw.qkv_w, (3 * d, d), torch.bfloat16).cpu()
        assert torch.equal(qkv, head_major((src["blocks.1.attn.qkv.weight"] * g1).to(torch.bfloat16), d))
        qb = dev_tensor(lw.qkv_b, (3 * d,), torch.float32).cpu()
        assert torch.equal(qb, head_major(src["blocks.1.attn.qkv.bias"], d))
        gu = dev_tensor(lw.gateup_w, (ipad, 2, d), torch.bfloat16).cpu()
        assert torch.equal(gu[:i, 0], (src["blocks.1.mlp.gate_proj.weight"] * g2).to(torch.bfloat16))
        assert torch.equal(gu[:i, 1], (src["blocks.1.mlp.up_proj.weight"] * g2).to(torch.bfloat16))
        assert int(gu[i:].abs().sum()) == 0
        gb = dev_tensor(lw.gateup_b, (ipad, 2), torch.float32).cpu()
        assert torch.equal(gb[:i, 1], src["blocks.1.mlp.up_proj.bias"]) and float(gb[i:].abs().sum()) == 0
        dw = dev_tensor(lw.down_w, (d, ipad), torch.bfloat16).cpu()
        assert torch.equal(dw[:, :i], src["blocks.1.mlp.down_proj.weight"].to(torch.bfloat16)) and int(dw[:, i:].abs().sum()) == 0
        pw = dev_tensor(w.patch_w, (d, 1176), torch.bfloat16).cpu()
        assert torch.equal(pw, src["patch_embed.proj.weight"].reshape(d, -1).to(torch.bfloat16))
        fc2b = dev_tensor(w.merger_fc2_b, (cfg.out_hidden_size,), torch.float32).cpu()
        assert torch.equal(fc2b, src["merger.mlp.2.bias"])


@pytest.mark.parametrize("m,ipad,k", [(300, 256, 160), (1024, 3456, 1280)])
def test_gemm_swiglu(m, ipad, k):
    a, b = rnd((m, k), 12), rnd((2 * ipad, k), 13, 0.05)
    bias = rnd((2 * ipad,), 14, 0.1, torch.float32)
    out = torch.zeros(m, ipad, dtype=torch.bfloat16, device=DEV)
    run_gemm(a, b, _lib.EPI_SWIGLU, out, bias=bias, ldo=ipad)
    z = a.float() @ b.float().t() + bias
    ref = torch.nn.functional.silu(z[:, 0::2]) * z[:, 1::2]
    close(out, ref, 1e-2)


@pytest.mark.parametrize("epi", ["gelu", "bf16", "f32"])
def test_gemm_bias_epilogues(epi):
    m, n, k = 520, 640, 640
    a, b = rnd((m, k), 15), rnd((n, k), 16, 0.05)
    bias = rnd((n,), 17, 0.1, torch.float32)
    z = a.float() @ b.float().t() + bias
    if epi == "gelu":
        out = torch.zeros(m, n, dtype=torch.bfloat16, device=DEV)
        run_gemm(a, b, _lib.EPI_BIAS_GELU, out, bias=bias)
        close(out, torch.nn.functional.gelu(z), 1e-2)
    else:
        perm = torch.randperm(m, generator=torch.Generator().manual_seed(18)).to(torch.int32).to(DEV)
        dt = torch.bfloat16 if epi == "bf16" else torch.float32
        out = torch.zeros(m, n, dtype=dt, device=DEV)
        run_gemm(a, b, _lib.EPI_BIAS_BF16 if epi == "bf16" else _lib.EPI_BIAS_F32, out, bias=bias, row_map=perm)
        ref = torch.empty_like(z)
        ref[perm.long()] = z
        close(out, ref, 1e-2 if epi == "bf16" else 1e-3)


def test_gemm_rejects_bad_args():
    a, b = rnd((64, 60), 1), rnd((64, 60), 2)   # K % 8 != 0
    out = torch.zeros(64, 64, dtype=torch.float32, device=DEV)
    with pytest.raises(ValueError):
        run_gemm(a, b, _lib.EPI_STORE_F32, out)


@pytest.mark.parametrize("rows,dim", [(1000, 1280), (37, 160), (16, 5120)])
def test_rmsnorm(rows, dim):
    x = rnd((rows, dim), 20, 2.0, torch.float32)
    w = (1 + 0.1 * rnd((dim,), 21, 1.0, torch.float32))
    out = torch.zeros(rows, dim, dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().b200vit_rmsnorm(x.data_ptr(), w.data_ptr(), out.data_ptr(), rows, dim, 1e-6, _stream()), "rmsnorm")
    torch.cuda.synchronize()
    close(out, tower_ref.rmsnorm_ref(x.cpu(), w.cpu()), 1e-2)


@pytest.mark.parametrize("lens,heads", [([64, 64, 32, 16, 64], 2), ([1024, 1024], 4), ([100, 7, 2304, 64, 1], 2), ([60], 16),
                                        ([1024] * 5 + [300, 512], 16)])   # 384 work items on 148 persistent CTAs
def test_attention_varlen(lens, heads):
    m, d = sum(lens), heads * 80
    qkv = rnd((m, 3 * d), 30, 1.0)
    out = torch.zeros(m, d, dtype=torch.bfloat16, device=DEV)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    _lib.check(_lib.lib().b200vit_attention(qkv.data_ptr(), out.data_ptr(), cu.ctypes.data_as(C.POINTER(C.c_int32)),
                                            len(lens), heads, _stream()), "attention")
    torch.cuda.synchronize()
    q, k, v = qkv.float().cpu().reshape(m, 3, heads, 80).unbind(1)
    ref = tower_ref.varlen_attention_ref(q, k, v, cu, 80 ** -0.5)
    close(out, ref, 1e-2)


@pytest.mark.parametrize("lens", [[1024], [2304], [640, 1024, 384]])
def test_attention_growing_row_maximum(lens):
    """Full-attention kernel with the lazy row maximum: keys get larger from one 128-key block to the next (and a few
    rows have outlier queries), so rows DO outgrow the maximum they were using by more than 2^8 several times and the
    rescale of O in TMEM (tcgen05.ld -> scale -> tcgen05.st) and of the running sum is exercised; also a segment whose
    first blocks are tiny so that the late maximum dominates the result (HF :182-204 softmax in fp32)."""
    heads = 2
    m, d = sum(lens), heads * 80
    g = torch.Generator().manual_seed(77)
    q = torch.randn(m, heads, 80, generator=g)
    k = torch.randn(m, heads, 80, generator=g)
    v = torch.randn(m, heads, 80, generator=g)
    ramp = (1.0 + 3.0 * (torch.arange(m) // 128 % 6).float()).view(m, 1, 1)      # block b of 128 keys scaled by 1, 4, 7, ...
    k = k * ramp
    q[::7] *= 4.0                                                              # outlier rows: logit ranges of hundreds
    k[: 128 * 2] *= 0.05                                                       # the first two blocks hardly matter
    qkv = torch.stack([q, k, v], 1).reshape(m, 3 * d).to(torch.bfloat16).to(DEV)
    out = torch.zeros(m, d, dtype=torch.bfloat16, device=DEV)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    _lib.check(_lib.lib().b200vit_attention(qkv.data_ptr(), out.data_ptr(), cu.ctypes.data_as(C.POINTER(C.c_int32)),
                                            len(lens), heads, _stream()), "attention")
    torch.cuda.synchronize()
    qf, kf, vf = qkv.float().cpu().reshape(m, 3, heads, 80).unbind(1)
    ref = tower_ref.varlen_attention_ref(qf, kf, vf, cu, 80 ** -0.5)
    assert torch.isfinite(out.float()).all()
    close(out, ref, 1e-2)


def test_attention_empty_segment_list():
    qkv = rnd((64, 3 * 160), 31)
    out = torch.zeros(64, 160, dtype=torch.bfloat16, device=DEV)
    cu = np.zeros(1, dtype=np.int32)
    assert _lib.lib().b200vit_attention(qkv.data_ptr(), out.data_ptr(), cu.ctypes.data_as(C.POINTER(C.c_int32)), 0, 2, _stream()) == 0


def test_cast_to_bf16():
    x = rnd((1000, 1176), 40, 1.0, torch.float32)
    out = torch.zeros(1000, 1176, dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().b200vit_cast_to_bf16(x.data_ptr(), 0, out.data_ptr(), x.numel(), _stream()), "cast")
    torch.cuda.synchronize()
    assert torch.equal(out, x.to(torch.bfloat16))


# ------------------------------------------------------------------ overlay + patchify (bit-exact)
def _clip(t, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (t, h, w, 3), dtype=torch.uint8, generator=g)


def _layer(h, w):
    from PIL import Image, ImageDraw
    vip = Image.new("RGBA", (w, h), (0, 0, 0, 0))
    d = ImageDraw.Draw(vip)
    d.ellipse([(w // 3, h // 4), (2 * w // 3, 3 * h // 4)], fill=(0, 255, 0, 100))
    d.rectangle([(w // 8, h // 7), (7 * w // 8, 6 * h // 7)], outline=(255, 0, 0, 200), width=3)
    return np.array(vip)


def _ops_for(t, h, w):
    flows = [(0.0, 0.0), (3.0, -2.0), (-4.7, 5.2), (-0.7, -0.4), (40.5, 10.25), (-900.0, 3.0)]
    ops, ref_ops = [], []
    for i in range(t):
        if i % 4 == 3:
            o = vit.FrameOp(mode=_lib.FRAME_CIRCLE, cx=w // 2 + i, cy=h // 3 - i, r=min(h, w) // 20, rgba=(9, 200, 30, 120))
            ref_ops.append(dict(mode=2, cx=o.cx, cy=o.cy, r=o.r, rgba=o.rgba))
        elif i % 4 == 2 and i > 4:
            o = vit.FrameOp()
            ref_ops.append(dict(mode=0))
        else:
            fx, fy = flows[i % len(flows)]
            sx, zx = vit.shift_from_flow(fx, w)
            sy, zy = vit.shift_from_flow(fy, h)
            o = vit.FrameOp(mode=_lib.FRAME_LAYER, sx=sx, sy=sy, zx=zx, zy=zy)
            ref_ops.append(dict(mode=1, sx=sx, zx=zx, sy=sy, zy=zy))
        ops.append(o)
    return ops, ref_ops


@pytest.mark.parametrize("t,h,w", [(6, 56, 84), (5, 28, 56)])
def test_overlay_composite_and_patchify_bitexact(t, h, w):
    frames = _clip(t, h, w, 50)
    layer = _layer(h, w)
    ops, ref_ops = _ops_for(t, h, w)
    spec = vit.OverlaySpec.from_rgba(layer, ops)
    fr = frames.to(DEV)
    fc = _lib.Frames(fr.data_ptr(), t, h, w)
    oc = spec.to_c(t)
    comp = torch.zeros_like(fr)
    _lib.check(_lib.lib().b200vit_overlay_composite(C.byref(fc), C.byref(oc), comp.data_ptr(), _stream()), "composite")
    torch.cuda.synchronize()
    ref = overlay_ref.overlay_clip_ref(frames.numpy(), layer, ref_ops)
    assert np.array_equal(comp.cpu().numpy(), ref)
    pv_ref, grid = patchify_ref.patchify_ref(ref)
    out = torch.zeros(pv_ref.shape, dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().b200vit_overlay_patchify(C.byref(fc), C.byref(oc), 14, 2, 2, out.data_ptr(), _stream()), "patchify")
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), torch.from_numpy(pv_ref).to(torch.bfloat16))


def test_overlay_palette_and_box_kinds():
    t, h, w = 4, 56, 56
    frames = _clip(t, h, w, 51)
    fr = frames.to(DEV)
    fc = _lib.Frames(fr.data_ptr(), t, h, w)
    ops = [vit.FrameOp(mode=_lib.FRAME_LAYER, sx=i, sy=-i) for i in range(t)]
    ref_ops = [dict(mode=1, sx=i, sy=-i) for i in range(t)]
    # palette: two colours
    idx = np.zeros((h, w), np.uint8)
    idx[10:30, 5:25] = 1
    idx[20:50, 20:40] = 2
    pal = [(0, 0, 0, 0), (255, 0, 0, 200), (0, 0, 255, 90)]
    lay = np.zeros((h, w, 4), np.uint8)
    lay[idx == 1] = pal[1]
    lay[idx == 2] = pal[2]
    for spec, lay_ref in ((vit.OverlaySpec.from_palette(idx, pal, ops), lay),
                          (vit.OverlaySpec.from_box((8, 6, 40, 50), 4, (0, 255, 0, 210), ops),
                           overlay_ref.box_layer_ref(h, w, (8, 6, 40, 50), 4, (0, 255, 0, 210))),
                          (vit.OverlaySpec.from_box((13, 4, 14, 13), 5, (0, 255, 0, 210), ops),
                           overlay_ref.box_layer_ref(h, w, (13, 4, 14, 13), 5, (0, 255, 0, 210)))):
        oc = spec.to_c(t)
        comp = torch.zeros_like(fr)
        _lib.check(_lib.lib().b200vit_overlay_composite(C.byref(fc), C.byref(oc), comp.data_ptr(), _stream()), "composite")
        torch.cuda.synchronize()
        assert np.array_equal(comp.cpu().numpy(), overlay_ref.overlay_clip_ref(frames.numpy(), lay_ref, ref_ops))


def test_patchify_golden_fixture(golden_dir):
    import os
    z = np.load(os.path.join(golden_dir, "patchify_hf.npz"))
    fr = torch.from_numpy(z["frames"]).to(DEV)
    t, h, w, _ = fr.shape
    fc = _lib.Frames(fr.data_ptr(), t, h, w)
    out = torch.zeros(z["pixel_values"].shape, dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().b200vit_overlay_patchify(C.byref(fc), None, 14, 2, 2, out.data_ptr(), _stream()), "patchify")
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), torch.from_numpy(z["pixel_values"]).to(torch.bfloat16))


def test_overlay_long_clip_multiple_launch_windows():
    """More than 128 frames: the per-frame ops travel as kernel parameters in windows of 128 frames; odd T exercises
    the repeat-last-frame padding across the window boundary."""
    t, h, w = 131, 28, 56
    frames = _clip(t, h, w, 52)
    layer = _layer(h, w)
    ops = [vit.FrameOp(mode=_lib.FRAME_LAYER, sx=(i % 7) - 3, sy=(i % 5) - 2) if i % 3 else vit.FrameOp() for i in range(t)]
    ref_ops = [dict(mode=1, sx=(i % 7) - 3, sy=(i % 5) - 2) if i % 3 else dict(mode=0) for i in range(t)]
    spec = vit.OverlaySpec.from_rgba(layer, ops)
    fr = frames.to(DEV)
    fc = _lib.Frames(fr.data_ptr(), t, h, w)
    oc = spec.to_c(t)
    comp = torch.zeros_like(fr)
    _lib.check(_lib.lib().b200vit_overlay_composite(C.byref(fc), C.byref(oc), comp.data_ptr(), _stream()), "composite")
    torch.cuda.synchronize()
    ref = overlay_ref.overlay_clip_ref(frames.numpy(), layer, ref_ops)
    assert np.array_equal(comp.cpu().numpy(), ref)
    pv_ref, grid = patchify_ref.patchify_ref(ref)
    assert grid.tolist() == [[66, 2, 4]]
    out = torch.zeros(pv_ref.shape, dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().b200vit_overlay_patchify(C.byref(fc), C.byref(oc), 14, 2, 2, out.data_ptr(), _stream()), "patchify")
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), torch.from_numpy(pv_ref).to(torch.bfloat16))


def test_overlay_extreme_shifts_and_no_overlay():
    t, h, w = 4, 56, 56
    frames = _clip(t, h, w, 53)
    layer = _layer(h, w)
    shifts = [(0, 0), (w, 0), (-w - 3, 5), (17, -h + 1)]   # fully off-screen and partially visible layers
    ops = [vit.FrameOp(mode=_lib.FRAME_LAYER, sx=sx, sy=sy) for sx, sy in shifts]
    ref_ops = [dict(mode=1, sx=sx, sy=sy) for sx, sy in shifts]
    fr = frames.to(DEV)
    fc = _lib.Frames(fr.data_ptr(), t, h, w)
    oc = vit.OverlaySpec.from_rgba(layer, ops).to_c(t)
    comp = torch.zeros_like(fr)
    _lib.check(_lib.lib().b200vit_overlay_composite(C.byref(fc), C.byref(oc), comp.data_ptr(), _stream()), "composite")
    torch.cuda.synchronize()
    assert np.array_equal(comp.cpu().numpy(), overlay_ref.overlay_clip_ref(frames.numpy(), layer, ref_ops))
    # no overlay at all == plain processor output
    pv_ref, _ = patchify_ref.patchify_ref(frames.numpy())
    out = torch.zeros(pv_ref.shape, dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().b200vit_overlay_patchify(C.byref(fc), None, 14, 2, 2, out.data_ptr(), _stream()), "patchify")
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), torch.from_numpy(pv_ref).to(torch.bfloat16))


def test_overlay_rejects_bad_specs():
    t, h, w = 2, 28, 28
    fr = _clip(t, h, w, 54).to(DEV)
    fc = _lib.Frames(fr.data_ptr(), t, h, w)
    comp = torch.zeros_like(fr)
    bad = vit.OverlaySpec(kind=_lib.LAYER_NONE, ops=[vit.FrameOp(mode=_lib.FRAME_LAYER)] * t)     # layer op without a layer
    assert _lib.lib().b200vit_overlay_composite(C.byref(fc), C.byref(bad.to_c(t)), comp.data_ptr(), _stream()) == -1
    bad2 = vit.OverlaySpec(kind=_lib.LAYER_NONE, ops=[vit.FrameOp(mode=_lib.FRAME_CIRCLE, r=500)] * t)   # radius too large
    assert _lib.lib().b200vit_overlay_composite(C.byref(fc), C.byref(bad2.to_c(t)), comp.data_ptr(), _stream()) == -1
    out = torch.zeros(4, 1176, dtype=torch.bfloat16, device=DEV)
    fc_bad = _lib.Frames(fr.data_ptr(), t, 27, 28)                                                 # not a multiple of 28
    assert _lib.lib().b200vit_overlay_patchify(C.byref(fc_bad), None, 14, 2, 2, out.data_ptr(), _stream()) == -1


def test_cast_fp16_and_tail():
    x = rnd((7, 1176), 41, 1.0, torch.float16)   # 8232 elements: not a multiple of the 8-wide vector path per thread block
    out = torch.zeros(7, 1176, dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().b200vit_cast_to_bf16(x.data_ptr(), 1, out.data_ptr(), x.numel(), _stream()), "cast")
    torch.cuda.synchronize()
    assert torch.equal(out, x.to(torch.bfloat16))


def test_clock_probe_reports_a_plausible_sm_clock():
    buf = torch.zeros(2, dtype=torch.int64, device=DEV)
    _lib.check(_lib.lib().b200vit_clock_probe(buf.data_ptr(), 50000, _stream()), "probe")
    torch.cuda.synchronize()
    cycles, ns = [int(v) for v in buf.cpu()]
    assert 50000 <= ns < 5_000_000
    assert 500.0 < cycles / ns * 1e3 < 2500.0          # MHz
    assert _lib.lib().b200vit_clock_probe(None, 1000, _stream()) == -1


def test_c_abi_demo_runs_on_the_gpu(tmp_path):
    """The plain-C host program (examples/c_abi_demo.c) drives the device through the C ABI alone: an exact GEMM, then
    b200vit_pack_weights from host arrays + b200vit_forward from frames, reproducible bit for bit."""
    import subprocess
    from test_host_cpu import build_c_demo
    exe = build_c_demo(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches against the host loop" in r.stdout and "device calls ok" in r.stdout, r.stdout
