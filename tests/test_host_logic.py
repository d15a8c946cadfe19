"""Host-side logic (CPU): scheduler, shift, pack / ids, regional masks, image processor, argument checks,
parameter tables - against the oracle's statements of the same reference lines."""
import numpy as np
import pytest
import torch

from oracle import flux_oracle as O
from reptext_b200 import config, weights
from reptext_b200._pipeline_common import RepTextPipelineBase, calculate_shift, retrieve_timesteps
from reptext_b200.pipeline_utils import SyntheticTextEncoders, SyntheticVAE, VaeImageProcessor, randn_tensor
from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler


@pytest.mark.parametrize("steps,tokens", [(28, 4096), (30, 4096), (4, 256), (28, 9216), (1, 1024)])
def test_scheduler_matches_oracle_sigmas(steps, tokens):
    sch = FlowMatchEulerDiscreteScheduler()
    c = sch.config
    mu = calculate_shift(tokens, c.base_image_seq_len, c.max_image_seq_len, c.base_shift, c.max_shift)
    assert mu == O.calculate_shift(tokens, 256, 4096, 0.5, 1.15)
    ts, n = retrieve_timesteps(sch, steps, "cpu", None, np.linspace(1.0, 1 / steps, steps), mu=mu)
    ots, osg = O.make_sigmas(steps, tokens)
    assert n == steps and torch.equal(ts, ots) and torch.equal(sch.sigmas, osg)
    assert sch.order == 1 and sch.step_index is None
    assert sch.index_for_timestep(ts[min(2, steps - 1)]) == min(2, steps - 1)


def test_shift_endpoints_and_errors():
    assert calculate_shift(256, 256, 4096, 0.5, 1.15) == pytest.approx(0.5)
    assert calculate_shift(4096, 256, 4096, 0.5, 1.15) == pytest.approx(1.15)
    assert calculate_shift(9216, 256, 4096, 0.5, 1.15) == pytest.approx(2.0167, abs=1e-3)
    sch = FlowMatchEulerDiscreteScheduler()
    with pytest.raises(ValueError):
        sch.set_timesteps(sigmas=[1.0, 0.5])                      # dynamic shifting needs mu
    with pytest.raises(ValueError):
        retrieve_timesteps(sch, 2, "cpu", [1, 2], [1.0, 0.5], mu=1.0)


def test_pack_unpack_ids_match_oracle():
    z = torch.randn(2, 16, 24, 40)
    p = RepTextPipelineBase._pack_latents(z, 2, 16, 24, 40)
    assert torch.equal(p, O.pack_latents(z))
    assert torch.equal(RepTextPipelineBase._unpack_latents(p, 24 * 8, 40 * 8, 16), z)
    ids = RepTextPipelineBase._prepare_latent_image_ids(1, 24, 40, "cpu", torch.float32)
    assert torch.equal(ids, O.prepare_latent_image_ids(24, 40))
    assert ids.shape == (12 * 20, 3) and ids[:, 0].abs().sum() == 0 and ids[21, 1] == 1 and ids[21, 2] == 1


def test_regional_mask_matches_oracle():
    m = np.zeros((256, 192), np.uint8)
    m[40:120, 30:150] = 255
    base = RepTextPipelineBase.__new__(RepTextPipelineBase)
    got = base._regional_masks([m], "cpu", torch.float32)[0]
    assert got.shape == (1, 16 * 12, 1) and torch.equal(got, O.regional_mask(m))
    assert base._regional_masks(None, "cpu", torch.float32) == []


def test_image_processor_and_standins():
    from PIL import Image
    ip = VaeImageProcessor(vae_scale_factor=16)
    img = Image.fromarray((np.random.RandomState(0).rand(40, 60, 3) * 255).astype(np.uint8))
    x = ip.preprocess(img, height=64, width=96)
    assert x.shape == (1, 3, 64, 96) and x.min() >= -1 and x.max() <= 1
    mp = VaeImageProcessor(vae_scale_factor=16, do_normalize=False, do_binarize=True, do_convert_grayscale=True)
    mk = mp.preprocess(Image.fromarray(np.full((64, 96), 200, np.uint8)), height=64, width=96)
    assert mk.shape == (1, 1, 64, 96) and set(mk.unique().tolist()) == {1.0}
    back = ip.postprocess(x, output_type="pil")
    assert back[0].size == (96, 64)
    vae = SyntheticVAE(dtype=torch.float32, device="cpu", posterior_std=0.1)
    d = vae.encode(x).latent_dist
    g1, g2 = torch.Generator().manual_seed(1), torch.Generator().manual_seed(1)
    assert d.sample(g1).shape == (1, 16, 8, 12) and torch.equal(d.sample(g2), vae.encode(x).latent_dist.sample(torch.Generator().manual_seed(1)))
    assert vae.decode(d.mode(), return_dict=False)[0].shape == (1, 3, 64, 96)
    enc = SyntheticTextEncoders(64, 32, dtype=torch.float32, device="cpu")
    a, b = enc.encode(["x", "y"], 16), enc.encode(["x"], 16)
    assert a[0].shape == (2, 16, 64) and a[1].shape == (2, 32) and torch.equal(a[0][0], b[0][0])
    r = randn_tensor((2, 3), generator=[torch.Generator().manual_seed(1), torch.Generator().manual_seed(1)])
    assert torch.equal(r[0], r[1])


def test_check_inputs_raises_like_the_reference():
    base = RepTextPipelineBase.__new__(RepTextPipelineBase)
    ok = dict(prompt="a", prompt_2=None, height=256, width=256)
    base.check_inputs(**ok)
    for bad in (dict(ok, height=250), dict(ok, prompt=None), dict(ok, prompt=3), dict(ok, prompt_embeds=torch.zeros(1)),
                dict(ok, prompt=None, prompt_embeds=torch.zeros(1)), dict(ok, max_sequence_length=513),
                dict(ok, callback_on_step_end_tensor_inputs=["noise_pred"])):
        with pytest.raises(ValueError):
            base.check_inputs(**bad)


def test_parameter_tables_match_the_survey_sizes():
    # the RepText ControlNet is 6 double + 0 single blocks = 2.1411 B parameters (4.28 GB bf16)
    assert weights.num_params(config.REPTEXT_CONTROLNET, "controlnet") == pytest.approx(2.1411e9, rel=1e-3)
    assert weights.num_params(config.FLUX_DEV, "transformer") == pytest.approx(11.90e9, rel=5e-3)
    sd = weights.random_state_dict(config.TINY_CONTROLNET, "controlnet", seed=3)
    sd2 = weights.random_state_dict(config.TINY_CONTROLNET, "controlnet", seed=3)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd) and "controlnet_blocks.1.weight" in sd
    z = weights.random_state_dict(config.TINY_CONTROLNET, "controlnet", seed=3, zero_init=True)
    assert z["controlnet_blocks.0.weight"].abs().sum() == 0 and z["controlnet_x_embedder.weight"].abs().sum() == 0


def test_models_refuse_cpu_and_missing_library(monkeypatch):
    from reptext_b200 import _lib, models
    with pytest.raises((ValueError, RuntimeError)):
        models.FluxControlNetModel.random_init(config.TINY_CONTROLNET, dtype=torch.float32, device="cpu")
    with pytest.raises(ValueError):
        _lib.dtype_code(torch.float16)
    with pytest.raises(ValueError):
        _lib.ptr(torch.zeros(1))
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/librt_reptext.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_text_to_render_span_host_logic():
    """``get_text_to_render`` (pipeline_flux_controlnet.py:257-280): the span is where the quoted text's tokens - without the
    first one and the EOS - sit in the prompt's T5 ids; ``'...'`` before ``"..."``; no window / no quotes raise as upstream
    (ValueError / IndexError); without tokenizers the option is refused."""
    from types import SimpleNamespace
    from reptext_b200.pipeline_flux_controlnet import FluxControlNetPipeline as P
    from reptext_b200.pipeline_utils import SyntheticTokenizer

    class Enc:
        dtype = torch.float32

        def __call__(self, ids, output_hidden_states=False):
            h = ids.float()[..., None].expand(*ids.shape, 8)
            return _Out(h)

    class _Out(tuple):
        def __new__(cls, h):
            o = super().__new__(cls, (h,))
            o.pooler_output = h[:, 0]
            return o

    tok2 = SyntheticTokenizer("t5", 1000, 512)
    mine = SimpleNamespace(tokenizer=SyntheticTokenizer("clip", 1000, 77), tokenizer_2=tok2, text_encoder=Enc(),
                           text_encoder_2=Enc(), tokenizer_max_length=77, _execution_device=torch.device("cpu"))
    for name in ("_locate_text_to_render", "_get_t5_prompt_embeds", "_get_clip_prompt_embeds", "_encode_text", "_text_ids"):
        setattr(mine, name, getattr(RepTextPipelineBase, name).__get__(mine))
    prompt = "a sign that says ' hello big world ' and \" not this \" in the city"
    pe, po, ids, start, end = P.encode_prompt(mine, prompt, None, device="cpu", max_sequence_length=32, get_text_to_render=True)
    words = prompt.split()
    assert (start, end) == (words.index("hello"), words.index("hello") + 4)      # hello big world ' (the closing quote stays)
    row = tok2([prompt], padding="max_length", max_length=32, truncation=True).input_ids[0]
    assert torch.equal(row[start:end], tok2(["hello big world '"]).input_ids[0][:4])
    assert pe.shape == (1, 32, 8) and po.shape == (1, 8) and ids.shape == (32, 3)
    assert P.encode_prompt(mine, 'say " only double quotes here " now', None, device="cpu", max_sequence_length=32,
                           get_text_to_render=True)[3:] == (2, 7)
    assert len(P.encode_prompt(mine, prompt, None, device="cpu", max_sequence_length=32)) == 3
    with pytest.raises(ValueError, match="No match found"):
        P.encode_prompt(mine, "says 'hello big world'now", None, device="cpu", max_sequence_length=32, get_text_to_render=True)
    with pytest.raises(IndexError):
        P.encode_prompt(mine, "no quotes at all", None, device="cpu", max_sequence_length=32, get_text_to_render=True)
    with pytest.raises(ValueError, match="encoded here"):
        P.encode_prompt(mine, prompt, None, device="cpu", prompt_embeds=pe, pooled_prompt_embeds=po, get_text_to_render=True)
    mine.tokenizer_2 = None
    with pytest.raises(ValueError):
        P.encode_prompt(mine, prompt, None, device="cpu", max_sequence_length=32, get_text_to_render=True)


def test_get_timesteps_tail_of_the_schedule():
    """``get_timesteps`` (pipeline_flux_controlnet.py:474-484): strength 1 keeps every step, 0.5 the second half, 0 none;
    the scheduler's begin index follows."""
    from types import SimpleNamespace
    sch = FlowMatchEulerDiscreteScheduler()
    sch.set_timesteps(sigmas=np.linspace(1.0, 1 / 10, 10), mu=1.0)
    me = SimpleNamespace(scheduler=sch)
    for strength, first in ((1.0, 0), (0.5, 5), (0.26, 7), (0.0, 10), (1.7, 0)):
        ts, n = RepTextPipelineBase.get_timesteps(me, 10, strength, "cpu")
        assert n == 10 - first and torch.equal(ts, sch.timesteps[first:]) and sch.begin_index == first
