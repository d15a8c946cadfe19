"""Prompt-encoder path on the GPU, through the C-ABI: the kernels of csrc/text_kernels.cu against
plain fp32 torch, and the T5EncoderModel / CLIPTextModel drop-ins (reptext_b200/text_encoders.py) against the oracle
(oracle/text_oracle.py, fp32, the same bf16-rounded weights; the oracle itself is pinned against transformers in
tests/test_text_oracle.py).

Tolerances: single kernels 4e-3 rel-L2; whole encoders 1e-2 on the hidden states (north_star's bf16 bar).
"""
import pytest
import torch
import torch.nn.functional as F

from util import rel_l2

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(autouse=True)
def _no_tf32():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32 = old


@pytest.fixture(scope="module")
def ops():
    from reptext_b200 import ops as _ops, _lib
    _lib.lib()
    return _ops


def _rand(shape, seed, scale=1.0, dtype=BF):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda", dtype=torch.float32) * scale).to(dtype)


@pytest.mark.parametrize("D,center", [(4096, False), (768, True), (128, True), (136, False)])
def test_norm_rows(ops, D, center):
    x = _rand((3, 50, D), 1, 3.0) + 0.7
    w, b = _rand((D,), 2, 0.2) + 1, (_rand((D,), 3, 0.2) if center else None)
    out = ops.norm_rows(x, w, b, 1e-5, center)
    xf = x.float()
    if center:
        ref = F.layer_norm(xf, (D,), w.float(), b.float(), 1e-5)
    else:
        ref = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * w.float()
    assert rel_l2(out, ref) < 4e-3


@pytest.mark.parametrize("B,S,H,mode", [(2, 512, 4, "bias"), (1, 77, 12, "causal"), (2, 40, 2, "bias"), (1, 33, 2, "plain"),
                                        (1, 100, 3, "causal"), (1, 300, 2, "causal"), (2, 257, 3, "bias"), (1, 640, 2, "plain")])
def test_text_attention(ops, B, S, H, mode):
    D = H * 64
    qkv = _rand((B, S, 3 * D), 4, 0.5)
    bias = _rand((H, 2 * S - 1), 5, 1.0, torch.float32) if mode == "bias" else None
    scale = 1.0 if mode == "bias" else 0.125
    out = ops.text_attention(qkv, H, scale, rel_bias=bias, causal=(mode == "causal"))
    q, k, v = [t.float().view(B, S, H, 64).transpose(1, 2) for t in qkv.split(D, dim=-1)]
    s = scale * (q @ k.transpose(-1, -2))
    i = torch.arange(S, device="cuda")
    if bias is not None:
        s = s + bias[:, (i[None, :] - i[:, None]) + S - 1][None]
    if mode == "causal":
        s = s + torch.full((S, S), float("-inf"), device="cuda").triu(1)
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, S, D)
    assert rel_l2(out, ref) < 4e-3
    # the default is the tcgen05 / TMEM / TMA kernel; the warp-level MMA form (option text_attn_mma) and the CUDA-core
    # form (option text_attn_simt) of the same attention are kept for A/B
    from reptext_b200 import _lib
    for opt in ("text_attn_mma", "text_attn_simt"):
        _lib.set_option(opt, 1)
        try:
            out2 = ops.text_attention(qkv, H, scale, rel_bias=bias, causal=(mode == "causal"))
        finally:
            _lib.set_option(opt, 0)
        assert rel_l2(out2, ref) < 4e-3, opt


def test_glu_act_and_embedding(ops):
    x = _rand((2, 37, 2 * 256), 6, 2.0)
    ref = x[..., :256].float() * x[..., 256:].float()
    assert rel_l2(ops.glu_act(x, 256, 0), ref) < 4e-3
    y = _rand((2, 37, 384), 7, 2.0)
    ref = y.float() * torch.sigmoid(1.702 * y.float())
    assert rel_l2(ops.glu_act(y, 384, 1), ref) < 4e-3
    table, pos = _rand((1000, 128), 8), _rand((16, 128), 9)
    ids = torch.randint(0, 1000, (3, 16), generator=torch.Generator().manual_seed(10))
    assert torch.equal(ops.embedding(table, ids), table[ids.cuda()])
    want = (table[ids.cuda()].float() + pos.float()[None]).to(BF)
    assert torch.equal(ops.embedding(table, ids, pos), want)
    with pytest.raises(IndexError):
        ops.embedding(table, torch.tensor([[0, 1000]]))


def _t5(cfg_over, seed):
    from oracle import text_oracle as TO
    from reptext_b200 import text_encoders as TE
    cfg = dict(TO.T5_XXL_CONFIG, **cfg_over)
    sd = {k: v.to(BF).float() for k, v in TO.random_state_dict(TO.t5_param_shapes(cfg), seed=seed).items()}
    return TO, TE, cfg, sd


@pytest.mark.parametrize("S", [64, 512])
def test_t5_encoder_small_vs_oracle(S):
    TO, TE, cfg, sd = _t5(dict(vocab_size=1000, d_model=256, d_ff=512, num_layers=4, num_heads=4), 11)
    m = TE.T5EncoderModel(cfg, sd)
    ids = torch.randint(0, 1000, (2, S), generator=torch.Generator().manual_seed(12))
    with torch.no_grad():
        want = TO.t5_encoder({k: v.cuda() for k, v in sd.items()}, cfg, ids.cuda())
    out = m(ids.cuda(), output_hidden_states=False)[0]
    assert out.shape == (2, S, 256) and out.dtype == BF
    assert rel_l2(out, want) < 1e-2


def test_t5_xxl_width_two_layers_vs_oracle():
    """T5-v1.1-XXL's real widths (d_model 4096, 64 heads, d_ff 10240) at the pipeline's 512 tokens, two blocks deep."""
    TO, TE, cfg, sd = _t5(dict(num_layers=2, vocab_size=4096), 13)
    m = TE.T5EncoderModel(cfg, sd)
    ids = torch.randint(0, 4096, (1, 512), generator=torch.Generator().manual_seed(14))
    with torch.no_grad():
        want = TO.t5_encoder({k: v.cuda() for k, v in sd.items()}, cfg, ids.cuda())
    out = m(ids.cuda())[0]
    e = rel_l2(out, want)
    print(f"T5-XXL width, 2 blocks, 512 tokens: rel-L2 {e:.2e}")
    assert e < 1e-2


def test_t5_xxl_full_depth_vs_oracle_and_stock_torch_bf16():
    """T5-v1.1-XXL as the pipeline runs it (24 blocks, d_model 4096, 64 heads, d_ff 10240, the 32128-row vocabulary, 512
    tokens; 4.7 B random parameters drawn on the device in bf16): the CUDA encoder against the fp32 oracle reading the SAME
    bf16 weights, and against what stock torch does with the oracle in bf16 (the reference's own numerics on this GPU).
    Bars: <= 3e-2 against fp32 after 24 residual blocks, and never worse than 1.25 x stock torch bf16."""
    from oracle import text_oracle as TO
    from reptext_b200 import text_encoders as TE
    cfg = dict(TO.T5_XXL_CONFIG)
    sd = TE.random_weights(TE.t5_param_shapes(TE.T5_XXL_CONFIG), 21, "cuda")
    assert len(sd) == len(TO.t5_param_shapes(cfg)) and sum(v.numel() for v in sd.values()) > 4.5e9

    class F32View(dict):                     # one fp32 copy of one tensor at a time (the whole model would be 19 GB)
        def __getitem__(self, k):
            return dict.__getitem__(self, k).float()

    m = TE.T5EncoderModel(None, sd, device="cuda")
    ids = torch.randint(2, cfg["vocab_size"], (1, 512), generator=torch.Generator().manual_seed(22)).cuda()
    ids[:, 40:] = 0                          # a 39-token prompt + EOS, padded to 512 like the pipeline's tokenizer call
    ids[:, 39] = 1
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            want = TO.t5_encoder(F32View(sd), cfg, ids)
            stock = TO.t5_encoder(sd, cfg, ids)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    out = m(ids)[0]
    assert out.shape == (1, 512, 4096) and out.dtype == BF and stock.dtype == BF
    e, e_stock = rel_l2(out, want), rel_l2(stock, want)
    print(f"T5-XXL, all 24 blocks, 512 tokens: rel-L2 against fp32 {e:.2e}; stock torch bf16 {e_stock:.2e} (ratio {e / e_stock:.2f})")
    assert e < 3e-2 and e <= 1.25 * e_stock


def test_clip_l_full_vs_oracle():
    """CLIP ViT-L/14 text model at full size (12 blocks, 768 wide, 77 tokens): hidden states and pooler_output."""
    from oracle import text_oracle as TO
    from reptext_b200 import text_encoders as TE
    from reptext_b200.pipeline_utils import SyntheticTokenizer
    cfg = dict(TO.CLIP_L_CONFIG)
    sd = {k: v.to(BF).float() for k, v in TO.random_state_dict(TO.clip_param_shapes(cfg), seed=15).items()}
    m = TE.CLIPTextModel(cfg, sd)
    tok = SyntheticTokenizer("clip", cfg["vocab_size"], 77)
    ids = tok(["a street sign that reads 'مرحبا'", "short"], padding="max_length", max_length=77, truncation=True).input_ids
    with torch.no_grad():
        last, pooled = TO.clip_text({k: v.cuda() for k, v in sd.items()}, cfg, ids.cuda())
    out = m(ids.cuda(), output_hidden_states=False)
    e1, e2 = rel_l2(out.last_hidden_state, last), rel_l2(out.pooler_output, pooled)
    print(f"CLIP-L: last_hidden_state rel-L2 {e1:.2e}, pooler_output rel-L2 {e2:.2e}")
    assert out.pooler_output.shape == (2, 768) and e1 < 1e-2 and e2 < 1e-2


def test_pipeline_encode_prompt_with_real_encoders():
    """``encode_prompt`` through tokenizers + both encoder drop-ins (RepText/pipeline_flux_controlnet.py:349-456) against
    the oracle on the same ids: CLIP sees ``prompt``, T5 sees ``prompt_2``."""
    import test_pipeline_gpu as TP
    from oracle import text_oracle as TO
    from reptext_b200 import text_encoders as TE
    from reptext_b200.pipeline_utils import SyntheticTokenizer
    pipe, TR, CN, _, _ = TP._tiny_pipe(BF, TRname="SMALL128_TRANSFORMER", CNname="SMALL128_CONTROLNET")
    tcfg = dict(TO.T5_XXL_CONFIG, vocab_size=1000, d_model=TR["joint_attention_dim"], d_ff=512, num_layers=2,
                num_heads=TR["joint_attention_dim"] // 64)
    ccfg = dict(TO.CLIP_L_CONFIG, vocab_size=1000, hidden_size=TR["pooled_projection_dim"], intermediate_size=256,
                num_hidden_layers=2, num_attention_heads=TR["pooled_projection_dim"] // 64)
    if TR["joint_attention_dim"] % 64 or TR["pooled_projection_dim"] % 64:
        pytest.skip("the small config's text widths are not multiples of 64")
    tsd = {k: v.to(BF).float() for k, v in TO.random_state_dict(TO.t5_param_shapes(tcfg), seed=16).items()}
    csd = {k: v.to(BF).float() for k, v in TO.random_state_dict(TO.clip_param_shapes(ccfg), seed=17).items()}
    pipe.text_encoder, pipe.text_encoder_2 = TE.CLIPTextModel(ccfg, csd), TE.T5EncoderModel(tcfg, tsd)
    pipe.tokenizer, pipe.tokenizer_2 = SyntheticTokenizer("clip", 1000, 77), SyntheticTokenizer("t5", 1000, 512)
    pipe.tokenizer_max_length = 77
    pe, po, ids = pipe.encode_prompt(prompt="a sign", prompt_2="a street sign that reads 'مرحبا'",
                                     num_images_per_prompt=2, max_sequence_length=128)
    assert pe.shape == (2, 128, tcfg["d_model"]) and po.shape == (2, ccfg["hidden_size"]) and ids.shape == (128, 3)
    t_ids = pipe.tokenizer_2(["a street sign that reads 'مرحبا'"], padding="max_length", max_length=128, truncation=True).input_ids
    c_ids = pipe.tokenizer(["a sign"], padding="max_length", max_length=77, truncation=True).input_ids
    with torch.no_grad():
        want_pe = TO.t5_encoder({k: v.cuda() for k, v in tsd.items()}, tcfg, t_ids.cuda())
        _, want_po = TO.clip_text({k: v.cuda() for k, v in csd.items()}, ccfg, c_ids.cuda())
    assert rel_l2(pe[0:1], want_pe) < 1e-2 and torch.equal(pe[0], pe[1])
    assert rel_l2(po[0:1], want_po) < 1e-2 and torch.equal(po[0], po[1])
