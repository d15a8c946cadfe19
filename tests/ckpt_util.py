"""Writes a tiny pipeline repository in diffusers' directory layout (what ``black-forest-labs/FLUX.1-dev`` looks like on
disk, RepText/infer.py:27-33) from seeded random weights - with plain ``safetensors`` / ``json`` / ``tokenizers`` calls, NOT
with the package under test, so that ``from_pretrained`` is checked against the layout and not against its own writer."""
import json
import os

import torch
from safetensors.torch import save_file


def _cfg(path, name, body):
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, name), "w") as fh:
        json.dump(body, fh)


def write_clip_tokenizer(path, max_length=77):
    chars = "abcdefghijklmnopqrstuvwxyz"
    vocab = {}
    for c in chars:
        vocab[c] = len(vocab)
    for c in chars:
        vocab[c + "</w>"] = len(vocab)
    vocab["<|startoftext|>"] = len(vocab)
    vocab["<|endoftext|>"] = len(vocab)
    _cfg(path, "vocab.json", vocab)
    with open(os.path.join(path, "merges.txt"), "w") as fh:
        fh.write("#version: 0.2\n")
    _cfg(path, "tokenizer_config.json", {"tokenizer_class": "CLIPTokenizer", "model_max_length": max_length,
                                         "bos_token": "<|startoftext|>", "eos_token": "<|endoftext|>",
                                         "unk_token": "<|endoftext|>", "pad_token": "<|endoftext|>"})
    return len(vocab)


def write_t5_tokenizer(path, max_length=512):
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, processors
    os.makedirs(path, exist_ok=True)
    vocab = [("<pad>", 0.0), ("</s>", 0.0), ("<unk>", 0.0), ("▁", -1.0)]
    vocab += [(c, -2.0 - 0.01 * i) for i, c in enumerate("abcdefghijklmnopqrstuvwxyz'")]
    tk = Tokenizer(models.Unigram(vocab, unk_id=2))
    tk.pre_tokenizer = pre_tokenizers.Metaspace()
    tk.decoder = decoders.Metaspace()
    tk.post_processor = processors.TemplateProcessing(single="$A </s>", special_tokens=[("</s>", 1)])
    tk.save(os.path.join(path, "tokenizer.json"))
    _cfg(path, "tokenizer_config.json", {"tokenizer_class": "T5TokenizerFast", "model_max_length": max_length,
                                         "eos_token": "</s>", "unk_token": "<unk>", "pad_token": "<pad>"})
    return len(vocab)


def write_tiny_flux_repo(root, TR, vae_cfg, t5_cfg, clip_cfg, seeds=(100, 28, 16, 17), shard_transformer=True):
    """``root`` becomes a FLUX.1-dev-shaped pipeline directory with the tiny architectures given; returns the state dicts
    (bf16-representable fp32 CPU tensors) keyed by component."""
    from oracle import text_oracle as TO
    from oracle import vae_oracle as V
    from reptext_b200 import weights
    bf = lambda sd: {k: v.to(torch.bfloat16) for k, v in sd.items()}
    tr = bf(weights.random_state_dict(TR, "transformer", seed=seeds[0]))
    vae = bf(V.random_state_dict(vae_cfg, seed=seeds[1]))
    t5 = bf(TO.random_state_dict(TO.t5_param_shapes(t5_cfg), seed=seeds[2]))
    clip = bf(TO.random_state_dict(TO.clip_param_shapes(clip_cfg), seed=seeds[3]))
    _cfg(root, "model_index.json", {
        "_class_name": "FluxPipeline", "_diffusers_version": "0.30.0.dev0",
        "scheduler": ["diffusers", "FlowMatchEulerDiscreteScheduler"], "text_encoder": ["transformers", "CLIPTextModel"],
        "text_encoder_2": ["transformers", "T5EncoderModel"], "tokenizer": ["transformers", "CLIPTokenizer"],
        "tokenizer_2": ["transformers", "T5TokenizerFast"], "transformer": ["diffusers", "FluxTransformer2DModel"],
        "vae": ["diffusers", "AutoencoderKL"]})
    # transformer: sharded like the real one (three files + index); config.json without axes_dims_rope like FLUX.1-dev's
    d = os.path.join(root, "transformer")
    body = {k: v for k, v in TR.items() if k not in ("axes_dims_rope", "out_channels")}
    _cfg(d, "config.json", dict(body, _class_name="FluxTransformer2DModel", _diffusers_version="0.30.0.dev0"))
    keys = sorted(tr)
    if shard_transformer:
        n = 3
        wm = {}
        for i in range(n):
            fn = f"diffusion_pytorch_model-{i + 1:05d}-of-{n:05d}.safetensors"
            part = {k: tr[k].contiguous() for k in keys[i::n]}
            save_file(part, os.path.join(d, fn), metadata={"format": "pt"})
            wm.update({k: fn for k in part})
        _cfg(d, "diffusion_pytorch_model.safetensors.index.json", {"metadata": {"total_size": 0}, "weight_map": wm})
    else:
        save_file({k: tr[k].contiguous() for k in keys}, os.path.join(d, "diffusion_pytorch_model.safetensors"))
    d = os.path.join(root, "vae")
    _cfg(d, "config.json", dict({k: (list(v) if isinstance(v, tuple) else v) for k, v in vae_cfg.items()},
                                _class_name="AutoencoderKL", act_fn="silu", sample_size=1024,
                                down_block_types=["DownEncoderBlock2D"] * 4, up_block_types=["UpDecoderBlock2D"] * 4))
    save_file({k: v.contiguous() for k, v in vae.items()}, os.path.join(d, "diffusion_pytorch_model.safetensors"))
    d = os.path.join(root, "text_encoder")
    _cfg(d, "config.json", dict(clip_cfg, architectures=["CLIPTextModel"], model_type="clip_text_model", torch_dtype="bfloat16"))
    save_file({k: v.contiguous() for k, v in clip.items()}, os.path.join(d, "model.safetensors"))
    d = os.path.join(root, "text_encoder_2")
    _cfg(d, "config.json", dict(t5_cfg, architectures=["T5EncoderModel"], model_type="t5", torch_dtype="bfloat16"))
    t5_disk = dict(t5)
    t5_disk["encoder.embed_tokens.weight"] = t5_disk["shared.weight"]      # transformers stores the tied copy too
    save_file({k: v.contiguous().clone() for k, v in t5_disk.items()}, os.path.join(d, "model.safetensors"))
    _cfg(os.path.join(root, "scheduler"), "scheduler_config.json",
         {"_class_name": "FlowMatchEulerDiscreteScheduler", "_diffusers_version": "0.30.0.dev0", "base_image_seq_len": 256,
          "base_shift": 0.5, "max_image_seq_len": 4096, "max_shift": 1.15, "num_train_timesteps": 1000, "shift": 3.0,
          "use_dynamic_shifting": True})
    write_clip_tokenizer(os.path.join(root, "tokenizer"))
    write_t5_tokenizer(os.path.join(root, "tokenizer_2"))
    f32 = lambda sd: {k: v.float() for k, v in sd.items()}
    return dict(transformer=f32(tr), vae=f32(vae), text_encoder_2=f32(t5), text_encoder=f32(clip))
