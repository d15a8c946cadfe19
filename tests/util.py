"""Shared helpers for the tests: seeded synthetic inputs of the shapes SURVEY.md §8(d) names."""
import numpy as np
import torch

from reptext_b200 import config, weights


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def box_mask(height, width, box):
    m = np.zeros([height, width], dtype=np.uint8)
    y0, y1, x0, x1 = box
    m[y0:y1, x0:x1] = 255
    return m


def synth_inputs(tr_cfg, cn_cfg, height, width, T, seed=0, batch=1, n_lines=1, dtype=torch.float32):
    """latents, prompt embeds, pooled, packed control latents, regional masks, ids (CPU, `dtype`)."""
    from oracle import flux_oracle as O

    g = torch.Generator().manual_seed(seed)
    lh, lw = 2 * (height // 16), 2 * (width // 16)
    N = (lh // 2) * (lw // 2)
    lat = O.pack_latents(torch.randn(batch, 16, lh, lw, generator=g))
    pe = torch.randn(batch, T, tr_cfg["joint_attention_dim"], generator=g)
    pooled = torch.randn(batch, tr_cfg["pooled_projection_dim"], generator=g)
    ccond = cn_cfg["in_channels"] + cn_cfg["extra_condition_channels"]
    conds = [torch.randn(batch, N, ccond, generator=g) for _ in range(n_lines)]
    masks = []
    for li in range(n_lines):
        y0 = (height // 4) * (li + 1) - height // 8
        m = box_mask(height, width, (y0, y0 + height // 6, width // 5, width - width // 5))
        masks.append(O.regional_mask(m))
    img_ids = O.prepare_latent_image_ids(lh, lw)
    txt_ids = torch.zeros(T, 3)
    c = lambda t: t.to(dtype)
    return dict(latents=c(lat), prompt_embeds=c(pe), pooled=c(pooled), conds=[c(x) for x in conds],
                masks=[c(m) for m in masks], img_ids=c(img_ids), txt_ids=c(txt_ids), N=N)
