"""``from_pretrained`` / ``save_pretrained`` on the GPU (RepText/infer.py:27-33): models loaded from a directory in
diffusers' layout are the models built from the same tensors directly - bit for bit, weights and outputs - and the
pipeline assembled by ``FluxControlNetPipeline.from_pretrained(dir, controlnet=...)`` produces the image of the pipeline
assembled by hand."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _forward_inputs(TR, CN, seed=3):
    from util import synth_inputs
    x = synth_inputs(TR, CN, 256, 256, 128, seed=seed)
    c = lambda v: v.to("cuda", BF)
    kw = dict(hidden_states=c(x["latents"]), encoder_hidden_states=c(x["prompt_embeds"]),
              pooled_projections=c(x["pooled"]), timestep=torch.tensor([0.75], device="cuda", dtype=BF),
              guidance=torch.tensor([3.5], device="cuda"), img_ids=x["img_ids"].cuda(), txt_ids=x["txt_ids"].cuda())
    return kw, c(x["conds"][0]), c(x["masks"][0])


@pytest.mark.parametrize("shards", [False, True])
def test_runtime_models_round_trip_through_a_directory(tmp_path, shards):
    from reptext_b200 import checkpoint as ck
    from reptext_b200 import config, models
    TR, CN = config.SMALL128_TRANSFORMER, config.SMALL128_CONTROLNET
    tr = models.FluxTransformer2DModel.random_init(TR, seed=1)
    cn = models.FluxControlNetModel.random_init(CN, seed=2)
    limit = 256 * 1024 if shards else 10 * 2 ** 30
    tr.save_pretrained(str(tmp_path / "transformer"), max_shard_size=limit)
    cn.save_pretrained(str(tmp_path / "controlnet"), max_shard_size=limit)
    assert os.path.exists(tmp_path / "transformer" / "diffusion_pytorch_model.safetensors.index.json") == shards
    tr2 = models.FluxTransformer2DModel.from_pretrained(str(tmp_path), subfolder="transformer", torch_dtype=BF)
    cn2 = models.FluxControlNetModel.from_pretrained(str(tmp_path / "controlnet"), torch_dtype=BF)
    assert cn2.config.extra_condition_channels == 64 and tuple(tr2.config.axes_dims_rope) == (16, 56, 56)
    for a, b in ((tr, tr2), (cn, cn2)):
        sa, sb = a.state_dict(), b.state_dict()
        assert sorted(sa) == sorted(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)
    kw, cond, mask = _forward_inputs(TR, CN)
    for net in ((cn, tr), (cn2, tr2)):
        bl, _ = net[0](controlnet_cond=cond, conditioning_scale=0.8, regional_mask=mask, return_dict=False, **kw)
        out = net[1](controlnet_block_samples=bl, return_dict=False, **kw)[0]
        if net[0] is cn:
            want_bl, want = [b.clone() for b in bl], out.clone()
    assert all(torch.equal(a, b) for a, b in zip(bl, want_bl)) and torch.equal(out, want)
    # a directory of the other class, a missing tensor and a tensor of the wrong shape are refused
    with pytest.raises(ValueError, match="holds a FluxControlNetModel"):
        models.FluxTransformer2DModel.from_pretrained(str(tmp_path / "controlnet"))
    sd = ck.load_state_dict(str(tmp_path / "controlnet"))
    broken = dict(sd)
    broken.pop("controlnet_x_embedder.weight")
    ck.write_config(str(tmp_path / "bad"), dict(CN), "FluxControlNetModel")
    ck.save_state_dict(str(tmp_path / "bad"), broken)
    with pytest.raises(RuntimeError, match="missing keys"):
        models.FluxControlNetModel.from_pretrained(str(tmp_path / "bad"))
    ck.write_config(str(tmp_path / "bad2"), dict(CN, extra_condition_channels=4), "FluxControlNetModel")
    ck.save_state_dict(str(tmp_path / "bad2"), sd)
    with pytest.raises(RuntimeError, match="controlnet_x_embedder.weight has shape"):
        models.FluxControlNetModel.from_pretrained(str(tmp_path / "bad2"))


def test_pipeline_from_pretrained_matches_the_hand_built_pipeline(tmp_path):
    """The two lines of RepText/infer.py:30-33 against a FLUX.1-dev-shaped directory (tiny architectures): sharded
    transformer, VAE, CLIP and T5 encoders, scheduler config, tokenizers read by transformers from the directory."""
    import ckpt_util
    import test_pipeline_gpu as TP
    from oracle import text_oracle as TO
    from oracle import vae_oracle as V
    from reptext_b200 import config, models, text_encoders, vae, weights
    from reptext_b200._pipeline_common import _load_tokenizer
    from reptext_b200.pipeline_flux_controlnet import FluxControlNetPipeline
    from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler
    TR, CN = config.SMALL128_TRANSFORMER, config.SMALL128_CONTROLNET
    vcfg = dict(V.FLUX_VAE_CONFIG, block_out_channels=(64, 128, 256, 256))
    tcfg = dict(TO.T5_XXL_CONFIG, vocab_size=1000, d_model=TR["joint_attention_dim"], d_ff=512, num_layers=2,
                num_heads=TR["joint_attention_dim"] // 64)
    ccfg = dict(TO.CLIP_L_CONFIG, vocab_size=1000, hidden_size=TR["pooled_projection_dim"], intermediate_size=256,
                num_hidden_layers=2, num_attention_heads=TR["pooled_projection_dim"] // 64)
    root = str(tmp_path / "FLUX.1-dev")
    sds = ckpt_util.write_tiny_flux_repo(root, TR, vcfg, tcfg, ccfg)
    cn_sd = {k: v.to(BF).float() for k, v in weights.random_state_dict(CN, "controlnet", seed=101).items()}
    models.FluxControlNetModel(CN, cn_sd).save_pretrained(str(tmp_path / "RepText"))

    controlnet = models.FluxControlNetModel.from_pretrained(str(tmp_path / "RepText"), torch_dtype=BF)
    pipe = FluxControlNetPipeline.from_pretrained(root, controlnet=controlnet, torch_dtype=BF).to("cuda")
    assert type(pipe.vae) is vae.AutoencoderKL and type(pipe.text_encoder_2) is text_encoders.T5EncoderModel
    assert type(pipe.tokenizer).__name__.startswith("CLIPTokenizer") and pipe.tokenizer_max_length == 77
    assert pipe.vae_scale_factor == 16 and pipe.scheduler.config.max_shift == 1.15

    hand = FluxControlNetPipeline(
        FlowMatchEulerDiscreteScheduler(), vae.AutoencoderKL(vcfg, sds["vae"]),
        text_encoders.CLIPTextModel(ccfg, sds["text_encoder"]), _load_tokenizer(os.path.join(root, "tokenizer"), "CLIPTokenizer"),
        text_encoders.T5EncoderModel(tcfg, sds["text_encoder_2"]), _load_tokenizer(os.path.join(root, "tokenizer_2"), None),
        models.FluxTransformer2DModel(TR, sds["transformer"]), models.FluxControlNetModel(CN, cn_sd))
    # component by component, bit for bit where the path is deterministic: the same tensors were loaded
    for name in ("transformer", "controlnet"):
        a, b = getattr(pipe, name).state_dict(), getattr(hand, name).state_dict()
        assert sorted(a) == sorted(b) and all(torch.equal(a[k], b[k]) for k in a), name
    for name in ("vae", "text_encoder", "text_encoder_2"):
        a, b = getattr(pipe, name)._w, getattr(hand, name)._w
        assert sorted(a) == sorted(b) and all(torch.equal(a[k], b[k]) for k in a), name
    ids = pipe.tokenizer_2(["a street sign that reads 'abc'"], padding="max_length", max_length=64, truncation=True,
                           return_tensors="pt").input_ids
    assert torch.equal(ids, hand.tokenizer_2(["a street sign that reads 'abc'"], padding="max_length", max_length=64,
                                             truncation=True, return_tensors="pt").input_ids)
    pe = [p.encode_prompt(prompt="a street sign", prompt_2="a street sign that reads 'abc'", max_sequence_length=64)
          for p in (pipe, hand)]
    assert pe[0][0].shape == (1, 64, tcfg["d_model"]) and torch.equal(pe[0][0], pe[1][0]) and torch.equal(pe[0][1], pe[1][1])
    # and the two lines of infer.py end to end.  (The VAE's GroupNorm statistics are accumulated with atomics, so two runs
    # of the SAME pipeline may differ in the last bf16 bit; the image is held to the run-to-run distance, not to equality.)
    H = W = 256
    _, cannys, poss, masks = TP._glyph_inputs(H, W, 2)
    outs = []
    for p in (pipe, hand, pipe):
        torch.cuda.manual_seed(29)           # the VAE posterior draws come from the global CUDA generator
        outs.append(p(prompt="a street sign", prompt_2="a street sign that reads 'abc'", height=H, width=W,
                      num_inference_steps=2, guidance_scale=3.5, control_image=cannys, control_position=poss,
                      control_mask=masks, controlnet_conditioning_scale=1.0, max_sequence_length=64,
                      generator=torch.Generator(device="cuda").manual_seed(5), output_type="pt").images)
    assert outs[0].shape == (1, 3, H, W) and torch.isfinite(outs[0]).all()
    from util import rel_l2
    print(f"from_pretrained vs hand-built {rel_l2(outs[0], outs[1]):.2e}, same pipeline twice {rel_l2(outs[0], outs[2]):.2e}")
    assert rel_l2(outs[0], outs[1]) < 3e-2
    # a keyword component replaces the directory's (diffusers' convention), here the scheduler
    sch = FlowMatchEulerDiscreteScheduler(max_shift=1.3)
    pipe2 = FluxControlNetPipeline.from_pretrained(root, controlnet=controlnet, scheduler=sch, vae=pipe.vae,
                                                   text_encoder=pipe.text_encoder, text_encoder_2=pipe.text_encoder_2,
                                                   transformer=pipe.transformer)
    assert pipe2.scheduler is sch and pipe2.transformer is pipe.transformer
    # examples/infer.py --base-model / --controlnet-model: the reference's script flow on the loaded pipeline
    import importlib.util
    spec = importlib.util.spec_from_file_location("infer_example", os.path.join(os.path.dirname(__file__), "..", "examples", "infer.py"))
    infer = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(infer)
    lat = infer.main(["--base-model", root, "--controlnet-model", str(tmp_path / "RepText"), "--steps", "2",
                      "--output-type", "latent", "--text", "مرحبا"])
    assert lat.shape == (1, 256, 64) and torch.isfinite(lat.float()).all()
