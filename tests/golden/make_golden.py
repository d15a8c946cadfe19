"""Generate the golden vectors in this directory from the CPU oracle (fp32).

    python tests/golden/make_golden.py

The reference ships no golden vectors (parity unpinned upstream), and its arithmetic lives in
`diffusers`, which cannot be installed here; these fixtures pin THIS repo's oracle so that a later
edit to it (or to torch) is noticed.  Inputs are regenerated from seeds by tests/util.py; only the
per-step packed latents are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import flux_oracle as O          # noqa: E402
from reptext_b200 import config, weights     # noqa: E402
from util import synth_inputs                 # noqa: E402

CASES = {
    # BASELINE.json configs[0]: 256x256, 4 Euler steps, batch 1, fp32
    "tiny_t2i": dict(kind="t2i", H=256, W=256, T=64, steps=4, lines=1),
    "tiny_t2i_2lines": dict(kind="t2i", H=256, W=256, T=64, steps=3, lines=2, cond_step=2),
    "tiny_inpaint": dict(kind="inpaint", H=256, W=256, T=64, steps=3, lines=1),
}


def run_case(name, spec):
    TR, CN, CNI = config.TINY_TRANSFORMER, config.TINY_CONTROLNET, config.TINY_INPAINT_CONTROLNET
    tr = weights.random_state_dict(TR, "transformer", seed=100)
    cn = weights.random_state_dict(CN, "controlnet", seed=101)
    x = synth_inputs(TR, CN, spec["H"], spec["W"], spec["T"], seed=102, n_lines=spec["lines"])
    ts, sg = O.make_sigmas(spec["steps"], x["N"])
    taps = []
    cb = lambda i, t, lat: taps.append(lat.clone())
    if spec["kind"] == "t2i":
        O.denoise_t2i(tr, TR, cn, CN, x["latents"], x["prompt_embeds"], x["pooled"], x["conds"], x["masks"],
                      x["txt_ids"], x["img_ids"], ts, sg, guidance_scale=3.5, conditioning_scale=1.0,
                      conditioning_step=spec.get("cond_step", 30), callback=cb)
    else:
        cni = weights.random_state_dict(CNI, "controlnet", seed=103)
        g = torch.Generator().manual_seed(104)
        neg_pe = torch.randn(1, spec["T"], TR["joint_attention_dim"], generator=g)
        neg_po = torch.randn(1, TR["pooled_projection_dim"], generator=g)
        cond_inp = torch.randn(1, x["N"], 68, generator=g)
        pe = torch.cat([neg_pe, x["prompt_embeds"]])
        po = torch.cat([neg_po, x["pooled"]])
        conds = [torch.cat([c] * 2) for c in x["conds"]]
        O.denoise_inpaint(tr, TR, cn, CN, cni, CNI, x["latents"], pe, po, conds, x["masks"],
                          torch.cat([cond_inp] * 2), x["txt_ids"], x["img_ids"], ts, sg, guidance_scale=3.5,
                          true_guidance_scale=3.5, conditioning_scale=1.0, conditioning_scale_inpaint=0.9, callback=cb)
    return torch.stack(taps).numpy()


def main():
    torch.set_num_threads(4)
    with torch.no_grad():
        for name, spec in CASES.items():
            arr = run_case(name, spec)
            np.savez_compressed(os.path.join(HERE, name + ".npz"), latents_per_step=arr.astype(np.float32))
            print(name, arr.shape, float(np.abs(arr).mean()))


if __name__ == "__main__":
    main()
