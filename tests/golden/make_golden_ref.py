"""Golden vectors made by RUNNING THE REFERENCE ITSELF: ``/root/reference/RepText/pipeline_flux_controlnet.py`` and
``pipeline_flux_controlnet_inpaint.py`` (with ``controlnet_flux.py``), imported by path and unmodified, over the stand-in
``diffusers`` of ``tests/ref_shim`` (Black-Forest-Labs block / autoencoder arithmetic from torchtitan, the real transformers
CLIP / T5 modules).  Cases and seeds: ``tests/ref_fixture.py``.

    python tests/golden/make_golden_ref.py            # needs /root/reference; the GPU box only reads the .npz files

Each ``ref_<case>.npz`` holds the packed latents after every step (the reference's own ``callback_on_step_end`` tap,
pipeline_flux_controlnet.py:1116-1123) and the tensors the reference's preparation code built before step 0 (packed
control latents, prompt embeddings, initial latents ...), recorded with forward pre-hooks on its modules.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import ref_fixture as F          # noqa: E402
import ref_run                   # noqa: E402


def run_case(name):
    case = F.CASES[name]
    pipe = F.reference_pipeline(case)
    box = F.capture_first_step(pipe)
    with torch.no_grad():
        taps, out = F.run_with_taps(pipe, F.call_kwargs(case))
    assert torch.equal(taps[-1], out)
    rec = F.prepared_from_capture(box)
    for k in ("block_samples0", "single_block_samples0"):       # large; test_reference_pin.py compares them live
        if k in rec:
            rec[k + "_norm"] = rec.pop(k).flatten(1).norm(dim=1)
    rec["latents_per_step"] = taps
    rec["timesteps"] = pipe.scheduler.timesteps.float().cpu()
    rec["sigmas"] = pipe.scheduler.sigmas.float().cpu()
    if case["dtype"] == "bf16":
        # fp32 truth for the bf16 cases: the oracle's loop (== the reference's to 1e-5, tests/test_reference_pin.py) in
        # fp32 on the SAME prepared tensors, timesteps rounded to bf16 like the bf16 run rounds them.  (Re-running the
        # reference's __call__ in fp32 would draw different noise: randn in bf16 is not rounded fp32 randn.)
        rec["latents_per_step_fp32"] = F.oracle_loop(name, rec, torch.float32, torch.bfloat16)
    return {k: v.numpy().astype(np.float32) for k, v in rec.items()}


def main():
    if not ref_run.available():
        raise SystemExit(f"the reference is not at {ref_run.REF}")
    torch.set_num_threads(8)
    for name in F.CASES:
        rec = run_case(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        print(name, rec["latents_per_step"].shape, float(np.abs(rec["latents_per_step"]).mean()),
              "%.1f KB" % (os.path.getsize(os.path.join(HERE, name + ".npz")) / 1024))


if __name__ == "__main__":
    main()
