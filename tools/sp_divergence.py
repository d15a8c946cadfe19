#!/usr/bin/env python
"""Where does the sequence-parallel (lock-step, world 8) forward leave the single-GPU forward?  Transformer only, full
width (D = 3072, 24 heads), S = 9728, growing block counts: prints rel-L2(lock step, single) per configuration."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import flux_oracle as O
from reptext_b200 import config, models
from reptext_b200.parallel import LockstepGroup, shard_tokens


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def main():
    dt, dev = torch.bfloat16, "cuda"
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    H5 = 1536
    N5, T = (H5 // 16) ** 2, 512
    g = torch.Generator(device=dev).manual_seed(15)
    r = lambda *s: torch.randn(*s, generator=g, device=dev).to(dt)
    y = dict(lat=r(1, N5, 64), pe=r(1, T, 4096), po=r(1, 768), img_ids=O.prepare_latent_image_ids(2 * (H5 // 16), 2 * (H5 // 16)).to(dev),
             txt_ids=torch.zeros(T, 3, device=dev), t=torch.tensor([0.62], device=dev, dtype=dt), g=torch.tensor([3.5], device=dev, dtype=dt))
    cfgs = [(1, 0), (0, 1), (2, 0), (0, 2), (4, 4), (19, 38)] if len(sys.argv) < 3 else [(1, 0), (0, 1)]
    for nl, ns in cfgs:
        cfg = dict(config.FLUX_DEV, num_layers=nl, num_single_layers=ns)
        tr = models.FluxTransformer2DModel.random_init(cfg, seed=100, dtype=dt, device=dev)
        kw = dict(hidden_states=y["lat"], encoder_hidden_states=y["pe"], pooled_projections=y["po"], timestep=y["t"], guidance=y["g"],
                  img_ids=y["img_ids"], txt_ids=y["txt_ids"])
        v = tr(return_dict=False, **kw)[0]
        v2 = tr(return_dict=False, **kw)[0]
        grp = LockstepGroup(world, device=dev)
        per_rank = [dict(hidden_states=shard_tokens(y["lat"], rk, world), encoder_hidden_states=shard_tokens(y["pe"], rk, world),
                         pooled_projections=y["po"], timestep=y["t"], guidance=y["g"], img_ids=shard_tokens(y["img_ids"], rk, world, dim=0),
                         txt_ids=shard_tokens(y["txt_ids"], rk, world, dim=0)) for rk in range(world)]
        v_sp = torch.cat(tr.forward_lockstep(grp, per_rank), dim=1)
        f = lambda t: t.float()
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        with torch.no_grad():
            sd = {k: t.float() for k, t in tr.state_dict().items()} if nl + ns <= 8 else None
            want = O.transformer_forward(sd, cfg, f(y["lat"]), f(y["pe"]), f(y["po"]), f(y["t"]), y["img_ids"], y["txt_ids"], f(y["g"]),
                                         None, None, dt) if sd is not None else None
        torch.backends.cuda.matmul.allow_tf32 = old
        msg = f"layers {nl:2d} + {ns:2d}: lock step vs single {rel(v_sp, v):.3e}   single vs single (repeat) {rel(v2, v):.1e}"
        if want is not None:
            msg += f"   single vs fp32 {rel(v, want):.3e}   lock step vs fp32 {rel(v_sp, want):.3e}"
        print(msg, flush=True)
        del tr, grp
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
