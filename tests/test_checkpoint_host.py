"""Host side of ``from_pretrained`` (RepText/infer.py:27-33): the diffusers / transformers directory layout is read and
written correctly, and every error the loaders can raise before touching a GPU is raised.  No CUDA call is made."""
import json
import os

import pytest
import torch

from reptext_b200 import checkpoint as ck
from reptext_b200 import config


def _sd(n=7, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {f"blocks.{i}.weight": torch.randn(8 + i, 16, generator=g).to(torch.bfloat16) for i in range(n)}


def test_single_file_and_sharded_round_trip(tmp_path):
    sd = _sd()
    files = ck.save_state_dict(str(tmp_path / "one"), sd)
    assert [os.path.basename(f) for f in files] == ["diffusion_pytorch_model.safetensors"]
    back = ck.load_state_dict(str(tmp_path / "one"))
    assert sorted(back) == sorted(sd) and all(torch.equal(back[k], sd[k]) and back[k].dtype == torch.bfloat16 for k in sd)
    # shards + diffusers' index file, transformers' stem
    files = ck.save_state_dict(str(tmp_path / "many"), sd, stem="model", max_shard_bytes=600)
    assert len(files) > 2 and all("-of-" in f for f in files)
    index = json.load(open(tmp_path / "many" / "model.safetensors.index.json"))
    assert sorted(index["weight_map"]) == sorted(sd)
    assert index["metadata"]["total_size"] == sum(v.numel() * 2 for v in sd.values())
    back = ck.load_state_dict(str(tmp_path / "many"))
    assert all(torch.equal(back[k], sd[k]) for k in sd)


def test_variant_and_legacy_bin(tmp_path):
    from safetensors.torch import save_file
    sd = _sd(3)
    d = tmp_path / "v"
    d.mkdir()
    save_file({k: v.contiguous() for k, v in sd.items()}, str(d / "diffusion_pytorch_model.fp16.safetensors"))
    with pytest.raises(OSError):
        ck.load_state_dict(str(d))
    assert sorted(ck.load_state_dict(str(d), variant="fp16")) == sorted(sd)
    d = tmp_path / "bin"
    d.mkdir()
    torch.save(sd, str(d / "diffusion_pytorch_model.bin"))
    assert all(torch.equal(ck.load_state_dict(str(d))[k], sd[k]) for k in sd)


def test_broken_shards_are_reported(tmp_path):
    sd = _sd()
    d = str(tmp_path / "s")
    files = ck.save_state_dict(d, sd, max_shard_bytes=600)
    os.remove(files[1])
    with pytest.raises(OSError, match="lists shards"):
        ck.load_state_dict(d)
    # a tensor named by the index that no shard holds
    d2 = str(tmp_path / "t")
    ck.save_state_dict(d2, sd, max_shard_bytes=600)
    p = os.path.join(d2, "diffusion_pytorch_model.safetensors.index.json")
    idx = json.load(open(p))
    idx["weight_map"]["ghost.weight"] = sorted(set(idx["weight_map"].values()))[0]
    json.dump(idx, open(p, "w"))
    with pytest.raises(OSError, match="no shard holds"):
        ck.load_state_dict(d2)


def test_config_and_directory_resolution(tmp_path):
    with pytest.raises(OSError, match="not a local directory"):
        ck.resolve_dir("black-forest-labs/FLUX.1-dev")          # a hub id: there is no hub client
    with pytest.raises(OSError):
        ck.resolve_dir(str(tmp_path), subfolder="transformer")
    ck.write_config(str(tmp_path / "m"), dict(config.REPTEXT_CONTROLNET), "FluxControlNetModel")
    cfg, cls = ck.read_config(str(tmp_path / "m"))
    assert cls == "FluxControlNetModel" and "_diffusers_version" not in cfg
    assert cfg["axes_dims_rope"] == [16, 56, 56] and cfg["extra_condition_channels"] == 64
    with pytest.raises(OSError, match="config.json"):
        ck.read_config(str(tmp_path))
    ck.write_model_index(str(tmp_path / "p"), "FluxPipeline", {"transformer": ("diffusers", "FluxTransformer2DModel")})
    assert ck.read_model_index(str(tmp_path / "p")) == {"transformer": ("diffusers", "FluxTransformer2DModel")}
    with pytest.raises(OSError, match="model_index.json"):
        ck.read_model_index(str(tmp_path / "m"))


def test_loaders_reject_the_wrong_directory_before_any_cuda_call(tmp_path):
    from reptext_b200.models import FluxControlNetModel, FluxTransformer2DModel
    from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler
    from reptext_b200.text_encoders import CLIPTextModel
    from reptext_b200.vae import AutoencoderKL
    for cls in (FluxControlNetModel, FluxTransformer2DModel, AutoencoderKL, CLIPTextModel, FlowMatchEulerDiscreteScheduler):
        with pytest.raises(OSError, match="not a local directory"):
            cls.from_pretrained("Shakker-Labs/RepText")
    d = str(tmp_path / "cn")
    ck.write_config(d, dict(config.TINY_CONTROLNET), "FluxControlNetModel")
    with pytest.raises(ValueError, match="holds a FluxControlNetModel"):
        FluxTransformer2DModel.from_pretrained(d)
    with pytest.raises(OSError, match="no weight file"):
        FluxControlNetModel.from_pretrained(d)
    v = str(tmp_path / "vae")
    ck.write_config(v, dict(use_quant_conv=True), "AutoencoderKL")
    with pytest.raises(ValueError, match="quant_conv"):
        AutoencoderKL.from_pretrained(v)


def test_scheduler_round_trip(tmp_path):
    from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler
    s = FlowMatchEulerDiscreteScheduler(max_shift=1.2)
    s.save_pretrained(str(tmp_path / "scheduler"))
    raw = json.load(open(tmp_path / "scheduler" / "scheduler_config.json"))
    assert raw["_class_name"] == "FlowMatchEulerDiscreteScheduler" and raw["max_shift"] == 1.2
    back = FlowMatchEulerDiscreteScheduler.from_pretrained(str(tmp_path), subfolder="scheduler")
    assert back.config.max_shift == 1.2 and back.config.base_image_seq_len == 256 and back.config.use_dynamic_shifting
    # a config re-saved by diffusers >= 0.31 carries every constructor default: accepted while they sit at the default
    raw.update(time_shift_type="exponential", stochastic_sampling=False, invert_sigmas=False, use_karras_sigmas=False,
               use_exponential_sigmas=False, use_beta_sigmas=False, shift_terminal=None)
    json.dump(raw, open(tmp_path / "scheduler" / "scheduler_config.json", "w"))
    assert FlowMatchEulerDiscreteScheduler.from_pretrained(str(tmp_path / "scheduler")).config.max_shift == 1.2
    raw["time_shift_type"] = "linear"
    json.dump(raw, open(tmp_path / "scheduler" / "scheduler_config.json", "w"))
    with pytest.raises(ValueError, match="time_shift_type"):
        FlowMatchEulerDiscreteScheduler.from_pretrained(str(tmp_path / "scheduler"))
    raw["time_shift_type"] = "exponential"
    raw["use_karras_sigmas"] = True
    json.dump(raw, open(tmp_path / "scheduler" / "scheduler_config.json", "w"))
    with pytest.raises(ValueError, match="use_karras_sigmas"):
        FlowMatchEulerDiscreteScheduler.from_pretrained(str(tmp_path / "scheduler"))


def test_tiny_repository_layout_and_tokenizers(tmp_path):
    """The fixture writer of the GPU test produces what the loaders expect; tokenizers load from the directory with
    transformers (upstream's, host-side) and never from the hub."""
    import ckpt_util
    from oracle import text_oracle as TO
    from oracle import vae_oracle as V
    from reptext_b200._pipeline_common import _load_tokenizer
    from reptext_b200.pipeline_flux_controlnet import FluxControlNetPipeline
    TR = config.SMALL128_TRANSFORMER
    vcfg = dict(V.FLUX_VAE_CONFIG, block_out_channels=(64, 128, 256, 256))
    tcfg = dict(TO.T5_XXL_CONFIG, vocab_size=1000, d_model=128, d_ff=512, num_layers=2, num_heads=2)
    ccfg = dict(TO.CLIP_L_CONFIG, vocab_size=1000, hidden_size=64, intermediate_size=256, num_hidden_layers=2,
                num_attention_heads=1)
    root = str(tmp_path / "flux")
    sds = ckpt_util.write_tiny_flux_repo(root, TR, vcfg, tcfg, ccfg)
    idx = ck.read_model_index(root)
    assert idx["tokenizer_2"] == ("transformers", "T5TokenizerFast") and idx["vae"][1] == "AutoencoderKL"
    tr = ck.load_state_dict(os.path.join(root, "transformer"))
    assert sorted(tr) == sorted(sds["transformer"]) and all(torch.equal(tr[k].float(), sds["transformer"][k]) for k in tr)
    cfg, cls = ck.read_config(os.path.join(root, "transformer"))
    assert cls == "FluxTransformer2DModel" and "axes_dims_rope" not in cfg       # like FLUX.1-dev's own config.json
    tok = _load_tokenizer(os.path.join(root, "tokenizer"), "CLIPTokenizer")
    ids = tok(["a sign"], padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids
    assert ids.shape == (1, 77) and tok.model_max_length == 77
    tok2 = _load_tokenizer(os.path.join(root, "tokenizer_2"), "T5TokenizerFast")
    ids2 = tok2(["a street sign"], padding="max_length", max_length=32, truncation=True, return_tensors="pt").input_ids
    assert ids2.shape == (1, 32) and int(ids2[0, -1]) == 0 and 1 in ids2[0].tolist()    # </s> then padding
    with pytest.raises(OSError):
        _load_tokenizer(os.path.join(root, "tokenizer_3"), None)
    # the pipeline loader names what the base repository does not hold, before loading anything
    with pytest.raises(ValueError, match="controlnet"):
        FluxControlNetPipeline.from_pretrained(root)
    with pytest.raises(TypeError, match="unexpected components"):
        FluxControlNetPipeline.from_pretrained(root, controlnet=object(), unet=object())
