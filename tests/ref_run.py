"""Import and run the REFERENCE's own files, unmodified, by path.  TEST INFRASTRUCTURE ONLY.

``/root/reference/RepText/{controlnet_flux,pipeline_flux_controlnet,pipeline_flux_controlnet_inpaint}.py`` import
``diffusers``, which cannot be installed in this container; ``tests/ref_shim`` supplies a stand-in package whose block
arithmetic is the Black-Forest-Labs code torchtitan ships (tests/ref_shim/README.md).  Nothing of the reference is copied:
the three files are executed where they lie.  The GPU box has no ``/root/reference``; there, ``available()`` is False and
only the committed outputs of these runs (``tests/golden/ref_*.npz``, made by ``tests/golden/make_golden_ref.py``) are used.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "ref_shim")
REF = os.environ.get("REPTEXT_REFERENCE", "/root/reference/RepText")
FILES = ("controlnet_flux.py", "pipeline_flux_controlnet.py", "pipeline_flux_controlnet_inpaint.py")

_cache = None


def shim_importable() -> bool:
    try:
        importlib.import_module("torchtitan.experiments.flux.model.layers")
        return True
    except Exception:
        return False


def available() -> bool:
    return all(os.path.isfile(os.path.join(REF, f)) for f in FILES) and shim_importable()


def shim():
    """The stand-in ``diffusers`` package (importable on the GPU box too: it only needs torchtitan)."""
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    return importlib.import_module("diffusers")


def _by_path(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """-> namespace(controlnet_flux, t2i, inpaint): the reference's three modules."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError(f"the reference is not at {REF} (or torchtitan's BFL modules are missing)")
    shim()
    old = sys.dont_write_bytecode
    sys.dont_write_bytecode = True                      # /root/reference is read-only
    try:
        cf = _by_path("controlnet_flux", os.path.join(REF, FILES[0]))      # the pipelines do `from controlnet_flux import`
        t2i = _by_path("reptext_ref_pipeline_flux_controlnet", os.path.join(REF, FILES[1]))
        inp = _by_path("reptext_ref_pipeline_flux_controlnet_inpaint", os.path.join(REF, FILES[2]))
    finally:
        sys.dont_write_bytecode = old
    _cache = SimpleNamespace(controlnet_flux=cf, t2i=t2i, inpaint=inp)
    return _cache
