"""Local checkpoints in the directory layout the reference loads with ``from_pretrained``.

``RepText/infer.py:27-33`` builds its models with::

    controlnet = FluxControlNetModel.from_pretrained("Shakker-Labs/RepText", torch_dtype=torch.bfloat16)
    pipe = FluxControlNetPipeline.from_pretrained("black-forest-labs/FLUX.1-dev", controlnet=controlnet, ...)

i.e. diffusers' layout: a model directory holds ``config.json`` and ``diffusion_pytorch_model.safetensors`` (or shards
listed in ``diffusion_pytorch_model.safetensors.index.json``); a pipeline directory holds ``model_index.json`` and one
sub-directory per component (``transformer/``, ``vae/``, ``text_encoder/``, ``text_encoder_2/``, ``tokenizer/``,
``tokenizer_2/``, ``scheduler/scheduler_config.json``); the text encoders use transformers' ``model.safetensors``.
This module reads and writes that layout for the drop-in classes of this package, so that a user of the reference
points the same two calls at a local copy of the same repositories.  There is no network on a serving box and no
hub client here: a name that is not a local directory raises ``OSError`` (what diffusers raises offline).

Host-side Python only: tensors are read with ``safetensors`` (``torch.load(weights_only=True)`` for ``.bin`` files),
handed to ``load_state_dict`` and from there to the C-ABI as device pointers.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, Optional, Sequence, Tuple

import torch

# stems diffusers / transformers use for weight files, in the order they are tried
DIFFUSERS_STEM = "diffusion_pytorch_model"
TRANSFORMERS_STEM = "model"
_LEGACY_BIN = {"diffusion_pytorch_model": "diffusion_pytorch_model.bin", "model": "pytorch_model.bin"}
_PRIVATE = ("_class_name", "_diffusers_version", "_name_or_path", "_commit_hash", "transformers_version",
            "architectures", "torch_dtype", "dtype")


def resolve_dir(name_or_path: str, subfolder: Optional[str] = None) -> str:
    """The local directory a ``from_pretrained`` call points at."""
    if not isinstance(name_or_path, (str, os.PathLike)):
        raise TypeError("pretrained_model_name_or_path must be a path to a local directory")
    path = os.fspath(name_or_path)
    if subfolder:
        path = os.path.join(path, subfolder)
    if not os.path.isdir(path):
        raise OSError(
            f"{path!r} is not a local directory.  reptext_b200 has no hub client (a serving box has no network): "
            "download the repository once (e.g. `huggingface-cli download <repo> --local-dir <dir>`) and pass <dir>.")
    return path


def read_config(dirpath: str, names: Sequence[str] = ("config.json",)) -> Tuple[dict, Optional[str]]:
    """``(config without the bookkeeping keys, _class_name or None)`` from the first of ``names`` that exists."""
    for n in names:
        p = os.path.join(dirpath, n)
        if os.path.isfile(p):
            with open(p) as fh:
                raw = json.load(fh)
            cls = raw.get("_class_name") or (raw.get("architectures") or [None])[0]
            return {k: v for k, v in raw.items() if k not in _PRIVATE}, cls
    raise OSError(f"no {' / '.join(names)} in {dirpath!r}")


def write_config(dirpath: str, config: dict, class_name: str, name: str = "config.json", transformers: bool = False) -> None:
    os.makedirs(dirpath, exist_ok=True)
    out = {"architectures": [class_name]} if transformers else {"_class_name": class_name, "_diffusers_version": "0.36.0"}
    for k, v in config.items():
        out[k] = list(v) if isinstance(v, tuple) else v
    with open(os.path.join(dirpath, name), "w") as fh:
        json.dump(out, fh, indent=2, sort_keys=True)
        fh.write("\n")


def _weight_files(dirpath: str, stem: str, variant: Optional[str]) -> Tuple[Sequence[str], Optional[Dict[str, str]]]:
    """Files holding the tensors of ``stem`` (one file, or the shards of an index) and the index's weight map."""
    v = f".{variant}" if variant else ""
    single = os.path.join(dirpath, f"{stem}{v}.safetensors")
    if os.path.isfile(single):
        return [single], None
    index = os.path.join(dirpath, f"{stem}.safetensors.index{v}.json")
    if not os.path.isfile(index):
        index = os.path.join(dirpath, f"{stem}{v}.safetensors.index.json")
    if os.path.isfile(index):
        with open(index) as fh:
            wm = json.load(fh)["weight_map"]
        files = sorted(set(wm.values()))
        missing = [f for f in files if not os.path.isfile(os.path.join(dirpath, f))]
        if missing:
            raise OSError(f"{index} lists shards that are not in {dirpath!r}: {missing[:3]}")
        return [os.path.join(dirpath, f) for f in files], wm
    legacy = os.path.join(dirpath, _LEGACY_BIN.get(stem, stem + ".bin"))
    if os.path.isfile(legacy):
        return [legacy], None
    return [], None


def load_state_dict(dirpath: str, stems: Iterable[str] = (DIFFUSERS_STEM, TRANSFORMERS_STEM),
                    variant: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """All tensors of the model stored in ``dirpath`` (CPU tensors, dtype as stored)."""
    for stem in stems:
        files, wm = _weight_files(dirpath, stem, variant)
        if not files:
            continue
        sd: Dict[str, torch.Tensor] = {}
        for f in files:
            if f.endswith(".safetensors"):
                from safetensors.torch import load_file
                part = load_file(f, device="cpu")
            else:
                part = torch.load(f, map_location="cpu", weights_only=True)
            dup = [k for k in part if k in sd]
            if dup:
                raise OSError(f"{f}: tensors {dup[:3]} are stored in more than one shard")
            sd.update(part)
        if wm is not None:
            lost = [k for k in wm if k not in sd]
            if lost:
                raise OSError(f"{dirpath!r}: the index names tensors no shard holds: {lost[:3]}")
        return sd
    raise OSError(f"no weight file ({' / '.join(s + '.safetensors' for s in stems)}, sharded index or .bin) in {dirpath!r}")


def save_state_dict(dirpath: str, sd: Dict[str, torch.Tensor], stem: str = DIFFUSERS_STEM,
                    max_shard_bytes: int = 10 * 2 ** 30) -> Sequence[str]:
    """Write ``sd`` as ``<stem>.safetensors`` - or as ``<stem>-0000i-of-0000n.safetensors`` shards with diffusers'
    index file when it exceeds ``max_shard_bytes`` (FLUX.1-dev's transformer ships as three 10 GB shards)."""
    from safetensors.torch import save_file
    os.makedirs(dirpath, exist_ok=True)
    items = [(k, v.detach().to("cpu").contiguous()) for k, v in sd.items()]
    shards, cur, size = [], {}, 0
    for k, t in items:
        n = t.numel() * t.element_size()
        if cur and size + n > max_shard_bytes:
            shards.append(cur)
            cur, size = {}, 0
        cur[k] = t
        size += n
    shards.append(cur)
    if len(shards) == 1:
        path = os.path.join(dirpath, f"{stem}.safetensors")
        save_file(shards[0], path, metadata={"format": "pt"})
        return [path]
    names, wm = [], {}
    for i, sh in enumerate(shards):
        fn = f"{stem}-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        save_file(sh, os.path.join(dirpath, fn), metadata={"format": "pt"})
        names.append(os.path.join(dirpath, fn))
        for k in sh:
            wm[k] = fn
    total = sum(t.numel() * t.element_size() for _, t in items)
    with open(os.path.join(dirpath, f"{stem}.safetensors.index.json"), "w") as fh:
        json.dump({"metadata": {"total_size": total}, "weight_map": wm}, fh, indent=2, sort_keys=True)
    return names


def read_model_index(dirpath: str) -> Dict[str, Tuple[Optional[str], Optional[str]]]:
    """``model_index.json`` of a pipeline directory: component name -> (library, class name)."""
    p = os.path.join(dirpath, "model_index.json")
    if not os.path.isfile(p):
        raise OSError(f"{dirpath!r} has no model_index.json: not a pipeline directory")
    with open(p) as fh:
        raw = json.load(fh)
    out = {}
    for k, v in raw.items():
        if k.startswith("_"):
            continue
        if isinstance(v, (list, tuple)) and len(v) == 2:
            out[k] = (v[0], v[1])
    return out


def write_model_index(dirpath: str, pipeline_class: str, components: Dict[str, Tuple[str, str]]) -> None:
    os.makedirs(dirpath, exist_ok=True)
    out = {"_class_name": pipeline_class, "_diffusers_version": "0.36.0"}
    for k, (lib, cls) in components.items():
        out[k] = [lib, cls]
    with open(os.path.join(dirpath, "model_index.json"), "w") as fh:
        json.dump(out, fh, indent=2, sort_keys=True)
        fh.write("\n")
