"""Model-level parity on the GPU through the C-ABI: FluxControlNetModel.forward and
FluxTransformer2DModel.forward against the oracle (oracle/flux_oracle.py) on identical random-init weights.

Bars (BASELINE.json north_star): <= 1e-4 rel-L2 against the fp32 oracle on the tiny config (fp32 kernels);
<= 1e-2 in bf16 (tcgen05 kernels), checked against the fp32 oracle evaluated on the same bf16-rounded weights.
"""
import pytest
import torch

from util import rel_l2, synth_inputs

pytestmark = pytest.mark.gpu


def _build(tr_cfg, cn_cfg, dtype, seed=100):
    from reptext_b200 import models, weights
    tr_sd = weights.random_state_dict(tr_cfg, "transformer", seed=seed)
    cn_sd = weights.random_state_dict(cn_cfg, "controlnet", seed=seed + 1)
    if dtype == torch.bfloat16:  # both sides see the same bf16-rounded parameters
        tr_sd = {k: v.to(dtype).float() for k, v in tr_sd.items()}
        cn_sd = {k: v.to(dtype).float() for k, v in cn_sd.items()}
    tr = models.FluxTransformer2DModel(tr_cfg, tr_sd, dtype=dtype)
    cn = models.FluxControlNetModel(cn_cfg, cn_sd, dtype=dtype)
    return tr, cn, tr_sd, cn_sd


def _oracle_time(v, dtype):
    """the reference rounds `timestep.to(dtype) * 1000` IN THE MODEL DTYPE
    (controlnet_flux.py:282-284).  The fp32 oracle is fed that rounded value so that both sides embed
    the same timestep (bf16(0.62) * 1000 is 620 in bf16, not 621.09)."""
    if dtype == torch.float32:
        return v
    return (v.to(dtype) * 1000).float() / 1000


def _oracle_on(device, sd):
    return {k: v.to(device) for k, v in sd.items()}


CASES = [
    # name, transformer cfg, controlnet cfg, H, W, T, dtype, tol
    ("tiny_fp32", "TINY_TRANSFORMER", "TINY_CONTROLNET", 256, 256, 64, torch.float32, 1e-4),
    ("tiny_fp32_ragged", "TINY_TRANSFORMER", "TINY_CONTROLNET", 208, 176, 40, torch.float32, 1e-4),
    ("small128_bf16", "SMALL128_TRANSFORMER", "SMALL128_CONTROLNET", 256, 256, 128, torch.bfloat16, 1e-2),
    ("small128_bf16_ragged", "SMALL128_TRANSFORMER", "SMALL128_CONTROLNET", 208, 176, 72, torch.bfloat16, 1e-2),
    ("tiny_bf16_simt", "TINY_TRANSFORMER", "TINY_CONTROLNET", 256, 256, 64, torch.bfloat16, 1e-2),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_controlnet_and_transformer_forward(case):
    from oracle import flux_oracle as O
    from reptext_b200 import config
    name, trn, cnn, H, Wd, T, dtype, tol = case
    TR, CN = getattr(config, trn), getattr(config, cnn)
    tr, cn, tr_sd, cn_sd = _build(TR, CN, dtype)
    x = synth_inputs(TR, CN, H, Wd, T, seed=7, batch=2, n_lines=2)
    if dtype == torch.bfloat16:
        x = {k: ([t.to(dtype).float() for t in v] if isinstance(v, list) else (v.to(dtype).float() if torch.is_tensor(v) else v))
             for k, v in x.items()}
    dev = "cuda"
    t = torch.tensor([0.62, 0.62])
    g = torch.tensor([3.5, 3.5])
    to, go = _oracle_time(t, dtype).to("cuda"), _oracle_time(g, dtype).to("cuda")
    # ---------------- oracle (fp32, on the GPU through stock torch: same function, faster than the CPU)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        osd_tr, osd_cn = _oracle_on(dev, tr_sd), _oracle_on(dev, cn_sd)
        xg = {k: ([t_.to(dev) for t_ in v] if isinstance(v, list) else (v.to(dev) if torch.is_tensor(v) else v))
              for k, v in x.items()}
        with torch.no_grad():
            ob, os_ = O.controlnet_forward(osd_cn, CN, xg["latents"], xg["conds"][0], 0.8, xg["prompt_embeds"],
                                           xg["pooled"], to, xg["img_ids"], xg["txt_ids"], go)
            onp = O.transformer_forward(osd_tr, TR, xg["latents"], xg["prompt_embeds"], xg["pooled"], to,
                                        xg["img_ids"], xg["txt_ids"], go, ob, None)
            onp0 = O.transformer_forward(osd_tr, TR, xg["latents"], xg["prompt_embeds"], xg["pooled"], to,
                                         xg["img_ids"], xg["txt_ids"], go, None, None)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert os_ is None
    # ---------------- CUDA path
    c = lambda v: v.to(dev, dtype)
    kw = dict(encoder_hidden_states=c(xg["prompt_embeds"]), pooled_projections=c(xg["pooled"]), timestep=c(t),
              img_ids=c(xg["img_ids"]), txt_ids=c(xg["txt_ids"]), guidance=g.to(dev))
    blocks, singles = cn(hidden_states=c(xg["latents"]), controlnet_cond=c(xg["conds"][0]), conditioning_scale=0.8,
                         return_dict=False, **kw)
    assert singles is None and len(blocks) == CN["num_layers"]
    for i, (got, want) in enumerate(zip(blocks, ob)):
        assert got.shape == want.shape and got.dtype == dtype
        assert rel_l2(got.float(), want) < tol, (name, "controlnet block", i, rel_l2(got.float(), want))
    # the oracle's samples go into the transformer, so that its error is measured on its own
    np0 = tr(hidden_states=c(xg["latents"]), return_dict=False, **kw)[0]
    assert rel_l2(np0.float(), onp0) < tol, (name, "transformer", rel_l2(np0.float(), onp0))
    np1 = tr(hidden_states=c(xg["latents"]), controlnet_block_samples=[c(s) for s in ob], return_dict=False, **kw)[0]
    assert rel_l2(np1.float(), onp) < tol, (name, "transformer + residuals", rel_l2(np1.float(), onp))
    assert rel_l2(onp, onp0) > 10 * tol  # the residual injection is visible at this tolerance
    # ---------------- fused regional mask + multi-line sum == the pipeline's host-side form
    m0, m1 = c(xg["masks"][0]), c(xg["masks"][1])
    b0, _ = cn(hidden_states=c(xg["latents"]), controlnet_cond=c(xg["conds"][0]), conditioning_scale=0.8,
               return_dict=False, regional_mask=m0, **kw)
    stacked = b0[0]._rt_stacked
    b01, _ = cn(hidden_states=c(xg["latents"]), controlnet_cond=c(xg["conds"][1]), conditioning_scale=0.8,
                return_dict=False, regional_mask=m1, accumulate_into=(stacked, None), **kw)
    with torch.no_grad():
        ob1, _ = O.controlnet_forward(osd_cn, CN, xg["latents"], xg["conds"][1], 0.8, xg["prompt_embeds"],
                                      xg["pooled"], to, xg["img_ids"], xg["txt_ids"], go)
    for i in range(len(ob)):
        want = xg["masks"][0] * ob[i] + xg["masks"][1] * ob1[i]
        assert rel_l2(b01[i].float(), want) < tol, (name, "mask+sum", i)


@pytest.mark.parametrize("dtype,trn,cnn,tol", [(torch.float32, "TINY_TRANSFORMER", "TINY_INPAINT_CONTROLNET", 1e-4),
                                                (torch.bfloat16, "SMALL128_TRANSFORMER", "SMALL128_CONTROLNET", 1e-2)])
def test_batch1_latents_broadcast_against_batch2_embeddings(dtype, trn, cnn, tol):
    """Inpaint pipeline true-CFG (pipeline_flux_controlnet_inpaint.py:1145): latents are NOT doubled."""
    from oracle import flux_oracle as O
    from reptext_b200 import config
    TR, CN = getattr(config, trn), getattr(config, cnn)
    tr, cn, tr_sd, cn_sd = _build(TR, CN, dtype)
    x = synth_inputs(TR, CN, 128, 192, 48, seed=9, batch=2, n_lines=1)
    dev = "cuda"
    r = (lambda v: v.to(dtype).float()) if dtype == torch.bfloat16 else (lambda v: v)
    lat1 = r(x["latents"][:1]).to(dev)
    pe, po = r(x["prompt_embeds"]).to(dev), r(x["pooled"]).to(dev)
    cond = r(x["conds"][0][:1]).to(dev)
    t, g = torch.tensor([0.3]).to(dev), torch.tensor([3.5]).to(dev)
    to, go = _oracle_time(t, dtype), _oracle_time(g, dtype)
    ii, ti = x["img_ids"].to(dev), x["txt_ids"].to(dev)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            ob, _ = O.controlnet_forward(_oracle_on(dev, cn_sd), CN, lat1, torch.cat([cond] * 2), 1.0, pe, po, to, ii, ti, go)
            onp = O.transformer_forward(_oracle_on(dev, tr_sd), TR, lat1, pe, po, to, ii, ti, go, ob, None)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    c = lambda v: v.to(dev, dtype)
    kw = dict(encoder_hidden_states=c(pe), pooled_projections=c(po), timestep=c(t), img_ids=ii, txt_ids=ti, guidance=g)
    bl, _ = cn(hidden_states=c(lat1), controlnet_cond=c(torch.cat([cond] * 2)), return_dict=False, **kw)
    for got, want in zip(bl, ob):
        assert got.shape[0] == 2 and rel_l2(got.float(), want) < tol
    got = tr(hidden_states=c(lat1), controlnet_block_samples=[c(s) for s in ob], return_dict=False, **kw)[0]
    assert got.shape == onp.shape and rel_l2(got.float(), onp) < tol


def test_model_argument_errors():
    from reptext_b200 import config, models
    cn = models.FluxControlNetModel.random_init(config.TINY_CONTROLNET, dtype=torch.float32)
    x = synth_inputs(config.TINY_TRANSFORMER, config.TINY_CONTROLNET, 64, 64, 8, batch=1)
    dev = "cuda"
    kw = dict(hidden_states=x["latents"].to(dev), controlnet_cond=x["conds"][0].to(dev),
              encoder_hidden_states=x["prompt_embeds"].to(dev), pooled_projections=x["pooled"].to(dev),
              timestep=torch.tensor([0.5], device=dev), img_ids=x["img_ids"].to(dev), txt_ids=x["txt_ids"].to(dev),
              guidance=torch.tensor([3.5], device=dev))
    cn(**kw)
    with pytest.raises(ValueError):   # controlnet_flux.py:297 raises ValueError on a mode/union mismatch
        cn(**dict(kw, controlnet_mode=torch.zeros(1, 1, dtype=torch.long, device=dev)))
    with pytest.raises(ValueError):
        cn(**dict(kw, guidance=None))
    with pytest.raises(ValueError):
        cn(**dict(kw, controlnet_cond=x["conds"][0][:, :, :100].to(dev)))
    with pytest.raises(ValueError):
        cn(**dict(kw, hidden_states=x["latents"]))  # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        models.FluxControlNetModel(config.TINY_CONTROLNET, {"x_embedder.weight": torch.zeros(1)}, dtype=torch.float32)


def test_one_denoise_step_is_cuda_graph_capturable():
    """A forward is a straight sequence of launches on the caller's stream (no allocation, no sync inside the library;
    the AdaLN side stream forks and joins inside the forward), so ControlNet + transformer + Euler capture into ONE
    CUDA graph; replaying it on new latents must reproduce the eager result bit for bit."""
    from reptext_b200 import config, models, ops
    dt, dev = torch.bfloat16, "cuda"
    TR, CN = config.SMALL128_TRANSFORMER, config.SMALL128_CONTROLNET
    tr = models.FluxTransformer2DModel.random_init(TR, seed=100, dtype=dt)
    cn = models.FluxControlNetModel.random_init(CN, seed=101, dtype=dt)
    x = synth_inputs(TR, CN, 256, 256, 128, seed=9)
    c = lambda v: v.to(dev, dt)
    lat, pe, po, cond, mask = c(x["latents"]), c(x["prompt_embeds"]), c(x["pooled"]), c(x["conds"][0]), c(x["masks"][0])
    ii, ti = x["img_ids"].to(dev), x["txt_ids"].to(dev)
    t, g = torch.tensor([0.62], device=dev, dtype=dt), torch.tensor([3.5], device=dev, dtype=dt)

    def step(z):
        kw = dict(hidden_states=z, encoder_hidden_states=pe, pooled_projections=po, timestep=t, guidance=g, img_ids=ii,
                  txt_ids=ti)
        bl, _ = cn(controlnet_cond=cond, conditioning_scale=0.9, regional_mask=mask, return_dict=False, **kw)
        v = tr(controlnet_block_samples=bl, return_dict=False, **kw)[0]
        return ops.euler_step(v, z, 0.62, 0.55)

    eager1 = step(lat)
    lat2 = torch.randn_like(lat)
    eager2 = step(lat2)
    static_in = lat.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step(static_in)                      # warm-up on the capture stream (workspace, kernel attributes)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        static_out = step(static_in)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out, eager1)
    static_in.copy_(lat2)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out, eager2)


@pytest.mark.parametrize("dtype,trn,cnn,H,Wd,T,tol", [
    (torch.float32, "TINY_TRANSFORMER", "TINY_CONTROLNET", 256, 192, 64, 1e-4),
    (torch.bfloat16, "SMALL128_TRANSFORMER", "SMALL128_CONTROLNET", 256, 256, 128, 1e-2)])
def test_controlnet_with_single_layers_and_single_sample_injection(dtype, trn, cnn, H, Wd, T, tol):
    """Row a7: a ControlNet WITH single-stream blocks (RepText/controlnet_flux.py:351-381: concatenated tokens through
    FluxSingleTransformerBlock, image rows collected; :390-392 zero-linears on them) and the transformer consuming
    ``controlnet_single_block_samples`` (diffusers: sample ``j // ceil(38 / n)`` added to the image rows after single
    block j).  The oracle's controlnet_forward is bit-identical to the reference's own forward for this configuration
    (tests/test_reference_pin.py::test_reference_controlnet_forward_equals_oracle[single_layers=2])."""
    from oracle import flux_oracle as O
    from reptext_b200 import config
    TR = getattr(config, trn)
    CN = dict(getattr(config, cnn), num_single_layers=2)
    tr, cn, tr_sd, cn_sd = _build(TR, CN, dtype)
    x = synth_inputs(TR, CN, H, Wd, T, seed=17, batch=2, n_lines=2)
    if dtype == torch.bfloat16:
        x = {k: ([t.to(dtype).float() for t in v] if isinstance(v, list) else (v.to(dtype).float() if torch.is_tensor(v) else v))
             for k, v in x.items()}
    dev = "cuda"
    t, g = torch.tensor([0.41, 0.41]), torch.tensor([3.5, 3.5])
    to, go = _oracle_time(t, dtype).to(dev), _oracle_time(g, dtype).to(dev)
    xg = {k: ([t_.to(dev) for t_ in v] if isinstance(v, list) else (v.to(dev) if torch.is_tensor(v) else v)) for k, v in x.items()}
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        osd_tr, osd_cn = _oracle_on(dev, tr_sd), _oracle_on(dev, cn_sd)
        with torch.no_grad():
            ob, os_ = O.controlnet_forward(osd_cn, CN, xg["latents"], xg["conds"][0], 0.8, xg["prompt_embeds"],
                                           xg["pooled"], to, xg["img_ids"], xg["txt_ids"], go)
            ob1, os1 = O.controlnet_forward(osd_cn, CN, xg["latents"], xg["conds"][1], 0.8, xg["prompt_embeds"],
                                            xg["pooled"], to, xg["img_ids"], xg["txt_ids"], go)
            onp = O.transformer_forward(osd_tr, TR, xg["latents"], xg["prompt_embeds"], xg["pooled"], to,
                                        xg["img_ids"], xg["txt_ids"], go, ob, os_)
            onp_b = O.transformer_forward(osd_tr, TR, xg["latents"], xg["prompt_embeds"], xg["pooled"], to,
                                          xg["img_ids"], xg["txt_ids"], go, ob, None)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert len(os_) == 2
    c = lambda v: v.to(dev, dtype)
    kw = dict(encoder_hidden_states=c(xg["prompt_embeds"]), pooled_projections=c(xg["pooled"]), timestep=c(t),
              img_ids=c(xg["img_ids"]), txt_ids=c(xg["txt_ids"]), guidance=g.to(dev))
    m0, m1 = c(xg["masks"][0]), c(xg["masks"][1])
    blocks, singles = cn(hidden_states=c(xg["latents"]), controlnet_cond=c(xg["conds"][0]), conditioning_scale=0.8,
                         return_dict=False, **kw)
    assert len(blocks) == CN["num_layers"] and len(singles) == 2
    for i, (got, want) in enumerate(zip(list(blocks) + list(singles), list(ob) + list(os_))):
        assert got.shape == want.shape and rel_l2(got.float(), want) < tol, (i, rel_l2(got.float(), want))
    # the transformer consumes both lists (the oracle's, so that its own error is what is measured)
    np1 = tr(hidden_states=c(xg["latents"]), controlnet_block_samples=[c(s) for s in ob],
             controlnet_single_block_samples=[c(s) for s in os_], return_dict=False, **kw)[0]
    assert rel_l2(np1.float(), onp) < tol, rel_l2(np1.float(), onp)
    assert rel_l2(onp, onp_b) > 10 * tol          # the single-sample injection is visible at this tolerance
    # mask + multi-line sum on the single samples too (pipeline_flux_controlnet.py:1065-1069, :1080-1087)
    b0, s0 = cn(hidden_states=c(xg["latents"]), controlnet_cond=c(xg["conds"][0]), conditioning_scale=0.8,
                return_dict=False, regional_mask=m0, **kw)
    b01, s01 = cn(hidden_states=c(xg["latents"]), controlnet_cond=c(xg["conds"][1]), conditioning_scale=0.8,
                  return_dict=False, regional_mask=m1, accumulate_into=(b0[0]._rt_stacked, s0[0]._rt_stacked), **kw)
    for i in range(2):
        want = xg["masks"][0] * os_[i] + xg["masks"][1] * os1[i]
        assert rel_l2(s01[i].float(), want) < tol, ("single mask+sum", i, rel_l2(s01[i].float(), want))


def test_unconsumed_controlnet_blocks_are_skipped_bit_identically():
    """with 6 ControlNet samples and a 19-block consumer only samples 0..4 are ever read.  Scaled down:
    3 ControlNet blocks feeding a 2-block transformer -> interval 1... use 4 CN blocks / 3 consumer blocks (interval 1,
    samples 0..2).  set_consumer() must (a) leave every consumed sample bit-identical, (b) return zeros for the rest,
    (c) launch fewer kernels, (d) give a bit-identical noise prediction."""
    from reptext_b200 import _lib, config, models
    TR = dict(config.SMALL128_TRANSFORMER, num_layers=3)
    CN = dict(config.SMALL128_CONTROLNET, num_layers=5)       # ceil(3 / 5) = 1 -> samples 0, 1, 2 are read; 3, 4 are not
    assert models.FluxControlNetModel.consumed_samples(6, 19) == 5          # the FLUX.1-dev + RepText pairing
    assert models.FluxControlNetModel.consumed_samples(5, 3) == 3
    assert models.FluxControlNetModel.consumed_samples(2, 4) == 2 and models.FluxControlNetModel.consumed_samples(0, 4) == 0
    dtype, dev = torch.bfloat16, "cuda"
    tr, cn, _, _ = _build(TR, CN, dtype)
    x = synth_inputs(TR, CN, 256, 256, 128, seed=23, batch=1)
    c = lambda v: v.to(dev, dtype)
    kw = dict(hidden_states=c(x["latents"]), encoder_hidden_states=c(x["prompt_embeds"]), pooled_projections=c(x["pooled"]),
              timestep=c(torch.tensor([0.5])), img_ids=c(x["img_ids"]), txt_ids=c(x["txt_ids"]),
              guidance=torch.tensor([3.5], device=dev))
    n0 = _lib.launch_count()
    full, _ = cn(controlnet_cond=c(x["conds"][0]), conditioning_scale=0.9, regional_mask=c(x["masks"][0]), return_dict=False, **kw)
    torch.cuda.synchronize()
    n_full = _lib.launch_count() - n0
    v_full = tr(controlnet_block_samples=full, return_dict=False, **kw)[0]
    cn.set_consumer(TR["num_layers"], TR["num_single_layers"])
    n0 = _lib.launch_count()
    live, _ = cn(controlnet_cond=c(x["conds"][0]), conditioning_scale=0.9, regional_mask=c(x["masks"][0]), return_dict=False, **kw)
    torch.cuda.synchronize()
    n_live = _lib.launch_count() - n0
    v_live = tr(controlnet_block_samples=live, return_dict=False, **kw)[0]
    for i in range(3):
        assert torch.equal(live[i], full[i]), i
    for i in (3, 4):
        assert float(full[i].float().abs().max()) > 0 and float(live[i].float().abs().max()) == 0.0
    assert n_live < n_full and torch.equal(v_live, v_full)
    cn.set_consumer(None)
    again, _ = cn(controlnet_cond=c(x["conds"][0]), conditioning_scale=0.9, regional_mask=c(x["masks"][0]), return_dict=False, **kw)
    assert torch.equal(again[4], full[4])


def test_step_invariant_cache_is_bit_identical_and_follows_its_inputs():
    """SURVEY.md 8f.2 (controlnet_flux.py:280-292, :316-317; pipeline_flux_controlnet.py:1029): with
    set_step_invariant_cache(True) the second forward on the same prompt tensors (a) launches fewer kernels, (b) returns
    the same bits as the uncached model at ANY timestep / latents, and (c) recomputes when a keyed tensor is replaced or
    modified in place (the key holds data_ptr and _version)."""
    from reptext_b200 import _lib, config
    TR, CN = config.SMALL128_TRANSFORMER, config.SMALL128_CONTROLNET
    dtype, dev = torch.bfloat16, "cuda"
    tr, cn, _, _ = _build(TR, CN, dtype)
    x = synth_inputs(TR, CN, 256, 256, 128, seed=31, batch=1)
    c = lambda v: v.to(dev, dtype)
    enc = c(x["prompt_embeds"])
    base = dict(encoder_hidden_states=enc, pooled_projections=c(x["pooled"]), img_ids=c(x["img_ids"]),
                txt_ids=c(x["txt_ids"]), guidance=torch.tensor([3.5], device=dev))

    def step(t, lat):
        kw = dict(base, hidden_states=lat, timestep=c(torch.tensor([t])))
        n0 = _lib.launch_count()
        blocks, _ = cn(controlnet_cond=c(x["conds"][0]), conditioning_scale=0.9, return_dict=False, **kw)
        v = tr(controlnet_block_samples=blocks, return_dict=False, **kw)[0]
        torch.cuda.synchronize()
        return [b.clone() for b in blocks], v.clone(), _lib.launch_count() - n0

    lat0, lat1 = c(x["latents"]), c(torch.randn_like(x["latents"]))
    ref0, ref1 = step(0.9, lat0), step(0.4, lat1)
    for net in (cn, tr):
        net.set_step_invariant_cache(True)
    got0, got1 = step(0.9, lat0), step(0.4, lat1)       # miss (fills the cache), then hit with other latents / timestep
    for ref, got in ((ref0, got0), (ref1, got1)):
        assert all(torch.equal(a, b) for a, b in zip(ref[0], got[0])) and torch.equal(ref[1], got[1])
    assert got1[2] < ref1[2], (got1[2], ref1[2])           # rope x2, guidance / pooled first linears, ... not relaunched
    # in-place change of a keyed tensor: _version moves, the cache must not be used
    enc.mul_(0.5)
    got2 = step(0.4, lat1)
    for net in (cn, tr):
        net.set_step_invariant_cache(False)
    ref2 = step(0.4, lat1)
    assert all(torch.equal(a, b) for a, b in zip(ref2[0], got2[0])) and torch.equal(ref2[1], got2[1])
    assert not torch.equal(ref2[1], ref1[1])
    # a replaced tensor (new storage) and a different token count
    for net in (cn, tr):
        net.set_step_invariant_cache(True)
    step(0.4, lat1)
    base["encoder_hidden_states"] = c(torch.randn_like(x["prompt_embeds"]))
    got3 = step(0.4, lat1)
    for net in (cn, tr):
        net.set_step_invariant_cache(False)
    ref3 = step(0.4, lat1)
    assert torch.equal(ref3[1], got3[1]) and all(torch.equal(a, b) for a, b in zip(ref3[0], got3[0]))


@pytest.mark.parametrize("batch", [1, 2])
def test_modulation_table_rows_are_the_bits_a_forward_computes(batch):
    """rt_model_build_modulation_table (controlnet_flux.py:282-291 + the AdaLN linears of every block, for all steps of
    an image in one pass): a forward that uses row i of the table returns the same bits as a forward that computes its
    modulation from (timestep_i, guidance, pooled) - ControlNet and transformer, batch 1 and the true-CFG batch 2 with a
    one-element timestep - launches fewer kernels, and ignores the timestep it is handed while a row is selected."""
    from reptext_b200 import _lib, config
    TR, CN = config.SMALL128_TRANSFORMER, config.SMALL128_CONTROLNET
    dtype, dev = torch.bfloat16, "cuda"
    tr, cn, _, _ = _build(TR, CN, dtype)
    x = synth_inputs(TR, CN, 256, 256, 128, seed=57, batch=batch)
    c = lambda v: v.to(dev, dtype)
    base = dict(encoder_hidden_states=c(x["prompt_embeds"]), pooled_projections=c(x["pooled"]), img_ids=c(x["img_ids"]),
                txt_ids=c(x["txt_ids"]), guidance=torch.tensor([3.5], device=dev))
    steps = torch.tensor([1.0, 0.8125, 0.53, 0.2578, 0.0371])
    lats = [c(torch.randn_like(x["latents"][:1])) for _ in steps]       # batch-1 latents broadcast when batch == 2

    def step(t, lat):
        kw = dict(base, hidden_states=lat, timestep=c(torch.tensor([t])))
        n0 = _lib.launch_count()
        blocks, _ = cn(controlnet_cond=c(x["conds"][0][:1]), conditioning_scale=0.9, return_dict=False, **kw)
        v = tr(controlnet_block_samples=blocks, return_dict=False, **kw)[0]
        torch.cuda.synchronize()
        return [b.clone() for b in blocks], v.clone(), _lib.launch_count() - n0

    ref = [step(float(t), lat) for t, lat in zip(steps, lats)]
    ts_all = c(steps)[:, None]                                          # [steps, 1]: broadcast over the batch
    for net in (cn, tr):
        assert net.build_modulation_table(ts_all, base["guidance"], base["pooled_projections"]) == len(steps)
    for i in (3, 0, 4, 1, 2):
        for net in (cn, tr):
            net.select_modulation(i)
        got = step(0.999 if i else 0.001, lats[i])                      # the timestep argument is not read
        assert all(torch.equal(a, b) for a, b in zip(ref[i][0], got[0])), i
        assert torch.equal(ref[i][1], got[1]), i
        assert got[2] < ref[i][2], (got[2], ref[i][2])
    for net in (cn, tr):
        net.select_modulation(None)
    back = step(float(steps[2]), lats[2])
    assert torch.equal(back[1], ref[2][1]) and back[2] == ref[2][2]
    with pytest.raises(ValueError):
        tr.select_modulation(len(steps))
