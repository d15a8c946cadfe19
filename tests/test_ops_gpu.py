"""Operator-level parity on the GPU, through the C-ABI: every kernel against the oracle's arithmetic
(oracle/flux_oracle.py for the Flux-specific pieces, plain fp32 torch for F.linear).

Tolerances: fp32 kernels 2e-5 rel-L2 (summation order differs); bf16 tensor-core kernels 4e-3 against an
fp32 evaluation of the same bf16 inputs (one bf16 output rounding is 2^-9 = 2e-3 relative).
"""
import math

import pytest
import torch
import torch.nn.functional as F

from util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from reptext_b200 import ops as _ops, _lib
    _lib.lib()
    return _ops


def _rand(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda", dtype=torch.float32) * scale).to(dtype)


def ref_epilogue(acc, mode, bias, out_prev, gate, extra, scale, mask, accumulate, norm_w, rope, hd, row0):
    """fp32 statement of the five epilogues of include/reptext_rt.h on acc [B, M, n]."""
    from reptext_b200 import _lib as L
    v = acc + (bias.float() if bias is not None else 0)
    if mode == L.EPI_BIAS:
        return v
    if mode == L.EPI_GELU:
        return F.gelu(v, approximate="tanh")
    if mode == L.EPI_GATE_RESID:
        g = gate[:, None, :] if gate is not None else 1.0
        r = out_prev.float() + g * v
        if extra is not None:
            r = r + extra.float()
        return r
    if mode == L.EPI_SCALE_MASK:
        r = v * scale
        if mask is not None:
            r = r * mask.float()[None, :, None]
        if accumulate:
            r = r + out_prev.float()
        return r
    if mode == L.EPI_QKNORM_ROPE:
        B, M, n = v.shape
        h = v.view(B, M, n // hd, hd)
        h = h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + 1e-6) * norm_w.float()
        if rope is not None:
            cs = rope[row0:row0 + M]                      # [M, hd/2, 2]
            cos, sin = cs[..., 0][None, :, None, :], cs[..., 1][None, :, None, :]
            x0, x1 = h[..., 0::2], h[..., 1::2]
            h = torch.stack([x0 * cos - x1 * sin, x1 * cos + x0 * sin], dim=-1).flatten(-2)
        return h.reshape(B, M, n)
    raise AssertionError(mode)


# ------------------------------------------------------------------------------------------------ GEMM
GEMM_CASES = [
    # name, dtype, impl, B, M, N-list(segments), K, modes
    ("simt_f32_ragged", torch.float32, 1, 2, 70, [50], 33, ["bias"]),
    ("simt_f32_gelu", torch.float32, 1, 1, 64, [96], 64, ["gelu"]),
    ("simt_f32_gate", torch.float32, 1, 2, 100, [64], 48, ["gate"]),
    ("simt_f32_mask", torch.float32, 1, 2, 100, [64], 48, ["mask"]),
    ("simt_f32_qk", torch.float32, 1, 2, 40, [128, 128, 128], 64, ["qk", "qk", "bias"]),
    ("simt_bf16", torch.bfloat16, 1, 1, 96, [64], 128, ["bias"]),
    ("tc1_bn64", torch.bfloat16, 2, 1, 128, [192], 128, ["bias"]),
    ("tc1_bn128", torch.bfloat16, 2, 2, 256, [384], 192, ["bias"]),
    ("tc1_bn256_ragged", torch.bfloat16, 2, 1, 200, [512], 320, ["gelu"]),
    ("tc1_k64", torch.bfloat16, 2, 1, 384, [256], 64, ["bias"]),
    ("tc1_kragged", torch.bfloat16, 2, 1, 128, [128], 72, ["bias"]),
    ("tc1_gate", torch.bfloat16, 2, 2, 300, [256], 512, ["gate"]),
    ("tc1_mask", torch.bfloat16, 2, 2, 256, [256], 256, ["mask"]),
    ("tc1_qkv", torch.bfloat16, 2, 2, 256, [256, 256, 256], 256, ["qk", "qk", "bias"]),
    ("tc1_qkvm", torch.bfloat16, 2, 1, 384, [256, 256, 256, 1024], 256, ["qk", "qk", "bias", "gelu"]),
    ("tc2_bn128", torch.bfloat16, 3, 2, 512, [384], 192, ["bias"]),
    ("tc2_bn256_ragged", torch.bfloat16, 3, 1, 300, [512], 320, ["gelu"]),
    ("tc2_gate", torch.bfloat16, 3, 2, 300, [256], 512, ["gate"]),
    ("tc2_qkv", torch.bfloat16, 3, 2, 256, [256, 256, 256], 256, ["qk", "qk", "bias"]),
    ("tc1_big", torch.bfloat16, 2, 1, 4096, [3072], 3072, ["bias"]),
    ("tc2_big", torch.bfloat16, 3, 1, 4096, [3072], 3072, ["bias"]),
    ("auto_big_k_long", torch.bfloat16, 0, 1, 1024, [3072], 15360, ["gate"]),
]


@pytest.mark.parametrize("case", GEMM_CASES, ids=[c[0] for c in GEMM_CASES])
def test_gemm(ops, case):
    from reptext_b200 import _lib as L
    name, dtype, impl, B, M, Ns, K, modes = case
    mode_map = dict(bias=L.EPI_BIAS, gelu=L.EPI_GELU, gate=L.EPI_GATE_RESID, mask=L.EPI_SCALE_MASK, qk=L.EPI_QKNORM_ROPE)
    hd = 128
    row0 = 8  # rows of the output / rope table are offset, like the image rows of the joint sequence
    A = _rand((B, M + 5, K), dtype, 1)
    Ntot = sum(Ns)
    out = _rand((B, row0 + M + 3, Ntot + 16), dtype, 2)
    out0 = out.clone()
    gate = _rand((B, Ntot), torch.float32, 3)
    extra = _rand((B, M, Ntot), dtype, 4)
    mask = torch.rand(M, device="cuda").to(dtype)
    rope_ang = torch.rand(row0 + M, hd // 2, device="cuda") * 6.28
    rope = torch.stack([rope_ang.cos(), rope_ang.sin()], dim=-1).contiguous()
    segs, refs, c0 = [], [], 8
    for i, (n, md) in enumerate(zip(Ns, modes)):
        W = _rand((n, K), dtype, 10 + i, K ** -0.5)
        bias = _rand((n,), dtype, 20 + i, 0.1)
        nw = (1 + 0.1 * _rand((hd,), torch.float32, 30 + i)).to(dtype)
        segs.append(ops.Segment(W=W, bias=bias, out=out, mode=mode_map[md], out_col0=c0,
                                norm_w=nw if md == "qk" else None))
        refs.append((W, bias, nw, c0, n, mode_map[md]))
        c0 += n
    p = ops.Problem(A=A, segs=segs, a_row0=2, m_rows=M, out_row0=row0, gate=gate, extra=extra, scale=0.7, mask=mask,
                    accumulate=True)
    ops.gemm([p], B, dtype, rope=rope, head_dim=hd, impl=impl)
    torch.cuda.synchronize()
    n_off = 0
    for (W, bias, nw, c0, n, md) in refs:
        acc = A[:, 2:2 + M].float() @ W.float().t()
        want = ref_epilogue(acc, md, bias, out0[:, row0:row0 + M, c0:c0 + n], gate[:, n_off:n_off + n],
                            extra[:, :, n_off:n_off + n], 0.7, mask, True, nw, rope, hd, row0)
        got = out[:, row0:row0 + M, c0:c0 + n].float()
        tol = 2e-5 if dtype == torch.float32 else 4e-3
        assert rel_l2(got, want) < tol, (name, md, rel_l2(got, want))
        n_off += n
    # nothing outside the addressed window may change
    keep = out.clone()
    keep[:, row0:row0 + M, 8:8 + Ntot] = out0[:, row0:row0 + M, 8:8 + Ntot]
    assert torch.equal(keep, out0), name


@pytest.mark.parametrize("band", [1, 2, 3, 5])
@pytest.mark.parametrize("case", [c for c in GEMM_CASES if c[0] in ("tc1_bn128", "tc1_gate", "tc2_gate", "tc1_qkvm", "tc2_big",
                                                                    "auto_big_k_long")], ids=lambda c: c[0])
def test_gemm_banded_tile_order(ops, case, band):
    """Option gemm_band: the row tiles in bands swept over all column tiles (the order long-K problems take on their
    own) - same results for every band height, batch, ragged last band, several segments."""
    from reptext_b200 import _lib as L
    L.set_option("gemm_band", band)
    try:
        test_gemm(ops, case)
        test_gemm_two_problems_joint_rows(ops, 3)
    finally:
        L.set_option("gemm_band", 0)


@pytest.mark.parametrize("bn", [224, 192, 160])
@pytest.mark.parametrize("impl", [2, 3], ids=["tc1", "tc2"])
@pytest.mark.parametrize("mode", ["bias", "gelu", "gate", "mask"])
def test_gemm_runtime_tile_width(ops, bn, impl, mode):
    """Option gemm_dyn_bn: single-segment launches on tiles of a run-time width (gemm_tc_kernel<256, cg, true>), the last
    column tile partial (768 = 3 x 224 + 96 = 4 x 160 + 128: whole and half 64-column store groups); against the fp32
    statement of the epilogue, the untouched window, and two problems with different weights in one launch."""
    from reptext_b200 import _lib as L
    L.set_option("gemm_dyn_bn", bn)
    try:
        test_gemm(ops, ("dyn", torch.bfloat16, impl, 2, 300, [768], 320, [mode]))
        test_gemm_two_problems_joint_rows(ops, impl)
    finally:
        L.set_option("gemm_dyn_bn", 0)


@pytest.mark.parametrize("shape", [(1216, 3072, 3072), (1152, 12288, 3072), (1216, 3072, 15360), (2432, 3072, 3072)],
                         ids=lambda s: "x".join(map(str, s)))
def test_gemm_auto_tile_width_on_shard_shapes(ops, shape):
    """The row counts of the sequence-parallel shards (1216 = 9728 / 8 rows per rank, 2432 at four ranks) leave the last
    wave of 256-wide tiles mostly empty; such launches pick a narrower run-time tile on their own.  Bit-identical to the
    256-wide kernel (an element's arithmetic does not depend on the tile that holds it), in place (gate + residual)."""
    from reptext_b200 import _lib as L
    dt = torch.bfloat16
    M, N, K = shape
    A, W = _rand((1, M, K), dt, 1), _rand((N, K), dt, 2, K ** -0.5)
    b, gate, res0 = _rand((N,), dt, 3, 0.1), _rand((1, N), torch.float32, 4), _rand((1, M, N), dt, 5)
    outs = []
    for opt in (0, -1, 224):
        L.set_option("gemm_dyn_bn", opt)
        out = res0.clone()
        ops.gemm([ops.Problem(A=A, segs=[ops.Segment(W=W, bias=b, out=out, mode=L.EPI_GATE_RESID)], gate=gate)], 1, dt)
        outs.append(out)
    L.set_option("gemm_dyn_bn", 0)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    want = res0[:, :256].float() + gate[:, None] * (A[:, :256].float() @ W.float().t() + b.float())
    assert rel_l2(outs[0][:, :256].float(), want) < 4e-3


def test_gemm_auto_band_on_the_long_k_shape(ops):
    """(4608, 3072, 15360): A is 141 MB, three waves - the launch bands itself (6 row tiles); bit-identical to the
    row-tiles-fastest order (the order changes which CTA computes a tile, not the arithmetic of a tile)."""
    from reptext_b200 import _lib as L
    dt = torch.bfloat16
    A, W = _rand((1, 4608, 15360), dt, 1), _rand((3072, 15360), dt, 2, 15360 ** -0.5)
    b = _rand((3072,), dt, 3, 0.1)
    outs = []
    for band in (0, -1, 4):
        L.set_option("gemm_band", band)
        out = torch.zeros(1, 4608, 3072, dtype=dt, device="cuda")
        ops.gemm([ops.Problem(A=A, segs=[ops.Segment(W=W, bias=b, out=out)])], 1, dt, impl=3)
        outs.append(out)
    L.set_option("gemm_band", 0)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert rel_l2(outs[0][:, :512].float(), A[:, :512].float() @ W.float().t() + b.float()) < 4e-3


@pytest.mark.parametrize("impl", [1, 2, 3], ids=["simt", "tc1", "tc2"])
def test_gemm_two_problems_joint_rows(ops, impl):
    """Text rows and image rows of one joint buffer, each with its own weights, in ONE launch; the
    image problem adds a ControlNet residual (`extra`) and the latents problem broadcasts batch 1."""
    from reptext_b200 import _lib as L
    dtype = torch.bfloat16
    B, T, Nimg, D = 2, 128, 384, 256
    x = _rand((B, T + Nimg, D), dtype, 1)
    res = _rand((B, T + Nimg, D), dtype, 2)
    res0 = res.clone()
    Wt, Wi = _rand((D, D), dtype, 3, D ** -0.5), _rand((D, D), dtype, 4, D ** -0.5)
    bt, bi = _rand((D,), dtype, 5, 0.1), _rand((D,), dtype, 6, 0.1)
    gt, gi = _rand((B, D), torch.float32, 7), _rand((B, D), torch.float32, 8)
    extra = _rand((B, Nimg, D), dtype, 9)
    pt = ops.Problem(A=x, segs=[ops.Segment(W=Wt, bias=bt, out=res, mode=L.EPI_GATE_RESID)], a_row0=0, m_rows=T,
                     out_row0=0, gate=gt)
    pi = ops.Problem(A=x, segs=[ops.Segment(W=Wi, bias=bi, out=res, mode=L.EPI_GATE_RESID)], a_row0=T, m_rows=Nimg,
                     out_row0=T, gate=gi, extra=extra)
    ops.gemm([pt, pi], B, dtype, impl=impl)
    want_t = res0[:, :T].float() + gt[:, None] * (x[:, :T].float() @ Wt.float().t() + bt.float())
    want_i = res0[:, T:].float() + gi[:, None] * (x[:, T:].float() @ Wi.float().t() + bi.float()) + extra.float()
    assert rel_l2(res[:, :T].float(), want_t) < 4e-3
    assert rel_l2(res[:, T:].float(), want_i) < 4e-3
    # broadcast A (batch-1 latents against batch-2 embeddings, inpaint pipeline :1145)
    lat = _rand((1, Nimg, 64), dtype, 10)
    We = _rand((D, 64), dtype, 11, 0.125)
    out = torch.zeros(B, Nimg, D, dtype=dtype, device="cuda")
    ops.gemm([ops.Problem(A=lat, segs=[ops.Segment(W=We, out=out)], broadcast_a=True)], B, dtype, impl=impl)
    want = (lat.float() @ We.float().t()).expand(B, -1, -1)
    assert rel_l2(out.float(), want) < 4e-3


def test_gemm_rejects_bad_arguments(ops):
    A = torch.zeros(1, 128, 60, dtype=torch.bfloat16, device="cuda")
    W = torch.zeros(128, 60, dtype=torch.bfloat16, device="cuda")
    out = torch.zeros(1, 128, 128, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):  # K % 8 != 0 cannot take the TMA path when forced
        ops.gemm([ops.Problem(A=A, segs=[ops.Segment(W=W, out=out)])], 1, torch.bfloat16, impl=2)
    ops.gemm([ops.Problem(A=A, segs=[ops.Segment(W=W, out=out)])], 1, torch.bfloat16, impl=0)  # auto -> SIMT
    with pytest.raises(ValueError):
        ops.gemm([], 1, torch.bfloat16)


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("dtype,hd,S,impl", [(torch.float32, 64, 320, 1), (torch.float32, 128, 100, 1),
                                             (torch.bfloat16, 128, 384, 1), (torch.bfloat16, 64, 77, 1)])
def test_attention_simt(ops, dtype, hd, S, impl):
    _attention_case(ops, dtype, hd, S, impl, 2, 3)


@pytest.mark.parametrize("S,B,H", [(256, 1, 1), (384, 2, 3), (320, 1, 2), (100, 1, 1), (4608, 1, 2), (1111, 2, 2)])
def test_attention_tcgen05(ops, S, B, H):
    _attention_case(ops, torch.bfloat16, 128, S, 2, B, H)


def test_attention_variants_are_not_in_the_product_library(ops):
    """The shipped library holds ONE attention kernel; the measured A/B forms (half-row, CTA pair, decoupled, ...) live
    in the -DRT_AB_VARIANTS build (python -m reptext_b200.build --ab) and are exercised by tools/attn_check.py."""
    from reptext_b200 import _lib as L
    if hasattr(L.lib(), "rt_debug_attn_trace"):
        pytest.skip("RT_LIB points at the A/B build")
    qkv = _rand((1, 256, 3 * 128), torch.bfloat16, 1)
    with pytest.raises(RuntimeError, match="product kernel only"):
        ops.attention(qkv, 1, 128, 0, 128, 256, impl=62)


def _attention_case(ops, dtype, hd, S, impl, B, H):
    qkv = _rand((B, S, 3 * H * hd + 8), dtype, 1)
    out = ops.attention(qkv, H, hd, 0, H * hd, 2 * H * hd, impl=impl)
    q, k, v = [qkv[:, :, i * H * hd:(i + 1) * H * hd].float().view(B, S, H, hd).transpose(1, 2) for i in range(3)]
    want = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, S, H * hd)
    tol = 2e-5 if dtype == torch.float32 else 4e-3
    assert rel_l2(out.float(), want) < tol


# ------------------------------------------------------------------------------------------------ HBM-bound kernels
@pytest.mark.parametrize("dtype,D", [(torch.float32, 256), (torch.bfloat16, 3072), (torch.bfloat16, 256),
                                     (torch.float32, 3072)])
def test_layernorm_modulate(ops, dtype, D):
    B, T, N = 2, 24, 100
    x = _rand((B, T + N, D), dtype, 1, 2.0) + 0.5
    mod = _rand((B, 4 * D), torch.float32, 2, 0.3)
    sh_t, sc_t, sh_i, sc_i = [mod[:, i * D:(i + 1) * D] for i in range(4)]
    out = ops.layernorm_modulate(x, [(0, T, sh_t, sc_t), (T, T + N, sh_i, sc_i)])
    xn = F.layer_norm(x.float(), (D,), None, None, 1e-6)
    want = torch.cat([xn[:, :T] * (1 + sc_t[:, None]) + sh_t[:, None], xn[:, T:] * (1 + sc_i[:, None]) + sh_i[:, None]], 1)
    assert rel_l2(out.float(), want) < (2e-6 if dtype == torch.float32 else 3e-3)


def test_rope_table_matches_oracle(ops):
    from oracle import flux_oracle as O
    ids = torch.cat([torch.zeros(16, 3), O.prepare_latent_image_ids(32, 48)]).cuda()
    tab = ops.rope_table(ids, (16, 56, 56))
    cos, sin = O.rope_table(ids.cpu(), (16, 56, 56))
    assert torch.allclose(tab[..., 0].cpu(), cos[:, 0::2], atol=2e-7)
    assert torch.allclose(tab[..., 1].cpu(), sin[:, 0::2], atol=2e-7)


@pytest.mark.parametrize("dtype,hd", [(torch.float32, 64), (torch.bfloat16, 128)])
def test_qknorm_rope_matches_oracle(ops, dtype, hd):
    from oracle import flux_oracle as O
    B, S, H = 2, 50, 3
    axes = (16, 24, 24) if hd == 64 else (16, 56, 56)
    ids = torch.cat([torch.zeros(10, 3), O.prepare_latent_image_ids(10, 16)]).cuda()
    tab = ops.rope_table(ids, axes)
    buf = _rand((B, S, H * hd), dtype, 1)
    w = (1 + 0.1 * _rand((hd,), torch.float32, 2)).to(dtype)
    want = O.apply_rope(O._rms(buf.cpu().float().view(B, S, H, hd), w.cpu().float()), O.rope_table(ids.cpu(), axes))
    ops.qknorm_rope_(buf, 0, H, hd, w, tab)
    assert rel_l2(buf.float().cpu().view(B, S, H, hd), want) < (2e-6 if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_euler_cfg_mask_blend_match_oracle(ops, dtype):
    from oracle import flux_oracle as O
    n = (2, 1000, 64)
    v, x = _rand(n, dtype, 1), _rand(n, dtype, 2)
    s0, s1 = 0.8731, 0.8012
    from reptext_b200 import _lib
    # The expectation is TORCH'S OWN evaluation of FlowMatchEulerDiscreteScheduler.step's expression
    #     prev = sample.to(float32) + (sigmas[i + 1] - sigmas[i]) * model_output ; prev.to(model_output.dtype)
    # with `sigmas` a float32 tensor ON THE DEVICE, which is where diffusers 0.36's set_timesteps(device=) leaves it: dt is
    # a 0-dim CUDA tensor, torch multiplies in model_output's dtype and rounds dt to it first.  The kernel follows that;
    # option euler_dt_host=1 follows the other form (sigmas on the CPU: dt stays fp32).
    for where, opt in (("cuda", 0), ("cpu", 1)):
        sig = torch.tensor([s0, s1], dtype=torch.float32, device=where)
        want = (x.to(torch.float32) + (sig[1] - sig[0]) * v).to(v.dtype)
        _lib.set_option("euler_dt_host", opt)
        try:
            got = ops.euler_step(v, x, s0, s1)
        finally:
            _lib.set_option("euler_dt_host", 0)
        assert torch.equal(got, want), (where, int((got != want).sum()))
    dt = (torch.tensor(s1) - torch.tensor(s0))
    dt = dt.to(dtype).float().item()                    # the device-sigmas form rounds dt to the model dtype
    got = ops.euler_step(v, x, s0, s1)
    if dtype == torch.float32:  # and the oracle's own function agrees in fp32
        assert torch.allclose(got.cpu(), O.euler_step(v.cpu(), torch.tensor(s0), torch.tensor(s1), x.cpu()), atol=1e-6)

    v2 = _rand((2, 1000, 64), dtype, 3)
    u, t = v2[0:1], v2[1:2]
    want = u + 3.5 * (t - u)
    assert torch.equal(ops.cfg_combine(v2, 3.5, False), want)
    assert torch.equal(ops.cfg_combine(v2, 3.5, True), t * 0.0)
    x1 = _rand((1, 1000, 64), dtype, 4)
    want_e = (x1.float() + (dt * want.float()).to(dtype).float()).to(dtype)
    assert torch.equal(ops.cfg_euler_step(v2, x1, 3.5, False, s0, s1), want_e)
    assert torch.equal(ops.cfg_euler_step(v2, x1, 3.5, True, s0, s1), x1)   # step 0 of the inpaint loop

    y = _rand((2, 96, 256), dtype, 5)
    m = torch.rand(1, 96, 1, device="cuda").to(dtype)
    acc = _rand((2, 96, 256), dtype, 6)
    assert torch.equal(ops.mask_scale_add(y, m, None, 1.0), m * y)
    assert torch.equal(ops.mask_scale_add(y, m, acc, 1.0), acc + m * y)
    assert torch.equal(ops.mask_scale_add(y, None, None, 0.5), y * 0.5)

    noise, z = _rand((1, 16, 32, 32), dtype, 7), _rand((1, 16, 32, 32), dtype, 8)
    gm = (torch.rand(1, 16, 32, 32, device="cuda") > 0.5)
    want = torch.where(gm, 0.10 * z + 1.00 * noise, noise)
    assert torch.equal(ops.glyph_init_blend(noise, z, gm.to(torch.uint8)), want)


def test_launch_counter_counts(ops):
    from reptext_b200 import _lib as L
    n0 = L.launch_count()
    x = torch.zeros(64, device="cuda")
    ops.euler_step(x, x, 1.0, 0.5)
    assert L.launch_count() == n0 + 1
