#!/usr/bin/env python
"""Benchmark of the RepText denoising step on B200 (BASELINE.json configs[1]: FLUX.1-dev-architecture
transformer + RepText ControlNet, random init, 1024x1024, bf16, batch 1 per GPU, synthetic data).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg2_small]

One "step" = one denoising step of the hot path: ControlNet forward (1 text line, regional mask) ->
transformer forward (19 double + 38 single blocks, residual injection) -> FlowMatch Euler step.
Rank 0 prints ONE JSON line (contract in the task statement).  Multi-GPU runs shard independent samples
(weak scaling, no data-path collective; the output latents are all-gathered over NCCL after the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "1024x1024 RepText denoise steps/s (ControlNet + FLUX.1-dev-arch transformer + Euler; 28 steps = 1 image)"
UNIT = "steps/s"
STEPS_PER_IMAGE = 28
PROMPT = "a street sign in city, with the text 'مرحبا بالعالم', filmfotos, film grain, reversal film photography"


def workload(name: str):
    from reptext_b200 import config
    if name == "cfg2":
        return dict(name="cfg2: FLUX.1-dev-arch (19+38 blocks, D=3072) + RepText ControlNet (6+0), 1024x1024, "
                         "N=4096 image + T=512 text tokens, 1 text line, bf16, batch 1 per GPU",
                    TR=config.FLUX_DEV, CN=config.REPTEXT_CONTROLNET, H=1024, W=1024, T=512)
    if name == "cfg5":  # BASELINE.json configs[4]: ONE 1536x1536 sample, tokens + attention heads sharded over all ranks
        return dict(name="cfg5: FLUX.1-dev-arch + RepText ControlNet, 1536x1536, N=9216 image + T=512 text tokens, "
                         "1 text line, bf16, batch 1; ONE sample sequence-parallel over all GPUs (heads sharded in "
                         "attention, exchanges fused into the GEMM / attention epilogues as NVLink peer stores)",
                    TR=config.FLUX_DEV, CN=config.REPTEXT_CONTROLNET, H=1536, W=1536, T=512, sp=True)
    if name == "cfg5_small":  # debugging aid: the sequence-parallel path on the 8-head test pair
        return dict(name="cfg5_small (debug): SP8 pair, 512x256", TR=config.SP8_TRANSFORMER, CN=config.SP8_CONTROLNET,
                    H=512, W=256, T=64, sp=True)
    if name == "cfg2_small":  # debugging aid only: same code path, 2+2 blocks, D=256
        return dict(name="cfg2_small (debug): 2+2 blocks, D=256, 256x256", TR=config.SMALL128_TRANSFORMER,
                    CN=config.SMALL128_CONTROLNET, H=256, W=256, T=128)
    raise SystemExit(f"unknown workload {name}")


def step_flops(TR, CN, N, T, lines=1, cn_live_layers=None):
    """2*M*N*K per GEMM, 4*S^2*D per attention; norms / softmax / elementwise not counted.
    ``cn_live_layers``: ControlNet double blocks actually run (the pipelines skip blocks whose sample the transformer
    never reads - 5 of 6 for FLUX.1-dev + RepText); None = all of them (the reference's count: 82.69 TFLOP at cfg 2)."""
    def model(c, kind):
        D = c["num_attention_heads"] * c["attention_head_dim"]
        S = N + T
        dbl = 24 * S * D * D + 4 * S * S * D
        nl = c["num_layers"] if (kind != "cn" or cn_live_layers is None) else cn_live_layers
        f = nl * dbl + c["num_single_layers"] * dbl
        f += 2 * N * D * c["in_channels"] + 2 * T * D * c["joint_attention_dim"]
        if kind == "cn":
            f += 2 * N * D * (c["in_channels"] + c["extra_condition_channels"])
            f += (nl + c["num_single_layers"]) * 2 * N * D * D
        else:
            f += 2 * N * D * c["out_channels"]
        return f
    return model(TR, "tr") + lines * model(CN, "cn")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a port of the reference path; diffusers itself is not installable here) on host cores
# ---------------------------------------------------------------------------------------------------------
_CPU_STATE = {}


def cpu_oracle_sample(wl):
    """ONE bounded sample of the step on the host cores, through the oracle (the reference path restated; the reference
    itself is Python over diffusers and cannot travel to this box): the reference's own ``FluxControlNetModel.forward`` at
    the full size (controlnet_flux.py:216-413: embedders, 6 double-stream blocks, 6 zero-linears, x scale) plus ONE
    single-stream block of the transformer, fp32, all host threads.  That is 9.6 of the step's 82.7 TFLOP; the step rate is
    the sample rate scaled by the FLOP ratio (both kinds of block run at the same FLOP rate on the CPU).
    -> (seconds for the sample, FLOPs of the sample, FLOPs of a step, description)."""
    from oracle import flux_oracle as O
    from reptext_b200 import weights
    TR, CN = wl["TR"], wl["CN"]
    N, T = (wl["H"] // 16) * (wl["W"] // 16), wl["T"]
    D = TR["num_attention_heads"] * TR["attention_head_dim"]
    st = _CPU_STATE
    if "cn_sd" not in st:
        g = torch.Generator().manual_seed(0)
        st["cn_sd"] = weights.random_state_dict(CN, "controlnet", seed=0)
        one = dict(TR, num_layers=0, num_single_layers=1)
        st["sg_sd"] = {k: v for k, v in weights.random_state_dict(one, "transformer", seed=1).items()
                       if k.startswith("single_transformer_blocks.0.")}
        st["lat"] = torch.randn(1, N, TR["in_channels"], generator=g)
        st["cond"] = torch.randn(1, N, CN["in_channels"] + CN["extra_condition_channels"], generator=g)
        st["pe"] = torch.randn(1, T, TR["joint_attention_dim"], generator=g)
        st["po"] = torch.randn(1, TR["pooled_projection_dim"], generator=g)
        st["x"], st["c"], st["temb"] = torch.randn(1, N, D, generator=g), torch.randn(1, T, D, generator=g), torch.randn(1, D, generator=g)
        st["img_ids"] = O.prepare_latent_image_ids(2 * (wl["H"] // 16), 2 * (wl["W"] // 16))
        st["txt_ids"] = torch.zeros(T, 3)
        st["rope"] = O.rope_table(torch.cat([st["txt_ids"], st["img_ids"]]), TR["axes_dims_rope"])
    with torch.no_grad():
        t0 = time.perf_counter()
        O.controlnet_forward(st["cn_sd"], CN, st["lat"], st["cond"], 1.0, st["pe"], st["po"], torch.tensor([0.7]),
                             st["img_ids"], st["txt_ids"], torch.tensor([3.5]))
        t1 = time.perf_counter()
        O.single_block(st["sg_sd"], "single_transformer_blocks.0.", st["x"], st["c"], st["temb"], st["rope"],
                       TR["num_attention_heads"])
        t2 = time.perf_counter()
    S = N + T
    f_block = 24 * S * D * D + 4 * S * S * D
    f_cn = step_flops(dict(TR, num_layers=0, num_single_layers=0, in_channels=0, joint_attention_dim=0, out_channels=0),
                      CN, N, T)
    f_sample = f_cn + f_block
    f_step = step_flops(TR, CN, N, T)
    desc = (f"oracle (fp32 torch CPU restatement of the reference path) - one full-size FluxControlNetModel.forward "
            f"({t1 - t0:.2f} s, {f_cn / 1e12:.2f} TFLOP) + one single-stream block ({t2 - t1:.2f} s, {f_block / 1e12:.2f} TFLOP) "
            f"at S={S}, D={D}; step = {f_step / 1e12:.2f} TFLOP, rate scaled by the FLOP ratio {f_step / f_sample:.2f}")
    return t2 - t0, f_sample, f_step, desc


def cpu_oracle_steps_per_s(wl, repeats=1):
    best = None
    for _ in range(repeats):
        sec, f_sample, f_step, desc = cpu_oracle_sample(wl)
        if best is None or sec < best[0]:
            best = (sec, f_sample, f_step, desc)
    sec, f_sample, f_step, desc = best
    return 1.0 / (sec * f_step / f_sample), desc


def run_reference(args, wl):
    """The reference arm: the reference's own CPU path for this step on the box's host cores.  The reference is Python over
    diffusers and cannot travel to (or be installed on) the GPU box, so the thing timed is the oracle - the restatement that
    tests/test_reference_pin.py pins to the reference's own files (bit-identical ControlNet forward).  Every "step" of this
    arm is ONE bounded sample (cpu_oracle_sample: 11.6 % of a step's FLOPs, ~4-6 s): --steps K --warmup W really runs K + W
    samples and `timed_region_s` is their wall time; `value` is the step rate those samples imply."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_oracle_sample(wl)
    secs, desc, f_sample, f_step = [], "", 1.0, 1.0
    t_begin = time.perf_counter()
    for _ in range(max(1, args.steps)):
        sec, f_sample, f_step, desc = cpu_oracle_sample(wl)
        secs.append(sec)
        if time.perf_counter() - t_begin > 170:      # keep the whole run within a few minutes
            break
    timed = time.perf_counter() - t_begin
    mean = sum(secs) / len(secs)
    v = 1.0 / (mean * f_step / f_sample)
    cores = torch.get_num_threads()
    sample = desc + f"; {len(secs)} samples timed in {timed:.1f} s, min {min(secs):.2f} / mean {mean:.2f} / max {max(secs):.2f} s"
    line = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=len(secs), warmup=args.warmup,
                # a "step" of this arm is ONE bounded sample (what was really run and timed: steps x ms_per_step = the
                # timed region); `value` is the whole-step rate those samples imply (sample time x FLOP ratio)
                ms_per_step=1000.0 * mean, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=dict(workload=wl["name"], note="CPU arm does not scale with --gpus; each timed "
                                              "step is one bounded sample (sample_fraction_of_step of a step's FLOPs), "
                                              "value = 1 / ms_per_whole_step_extrapolated",
                                              sample_fraction_of_step=f_sample / f_step, timed_region_s=timed,
                                              samples_timed=len(secs), ms_per_whole_step_extrapolated=1000.0 / v),
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                images_per_s=v / STEPS_PER_IMAGE)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU baseline leg: the SAME step through stock torch on the same B200 (BASELINE.md section 4 / SURVEY.md 2a: "the kernel
# to beat is torch 2.11's cuBLASLt / SDPA bf16 path running the oracle on the B200").  A baseline, like cpu_baseline: the
# oracle is the thing measured HERE, never the product.
# ---------------------------------------------------------------------------------------------------------
def gpu_baseline_leg(tr, cn, TR, CN, lat, pe, po, cond, mask, img_ids, txt_ids, sigmas, tsd, steps, ours_first):
    from oracle import flux_oracle as O
    dt = torch.bfloat16
    tr_sd, cn_sd = tr.state_dict(), cn.state_dict()      # the product's own device tensors (diffusers names), no copy
    g = torch.tensor([3.5], device=lat.device)
    sig = sigmas.to(lat.device)
    m3 = mask.reshape(1, -1, 1)

    def step(i, latents):
        j = i % STEPS_PER_IMAGE
        timestep = tsd[j].expand(1).to(dt)
        b, _ = O.controlnet_forward(cn_sd, CN, latents, cond, 1.0, pe, po, timestep / 1000, img_ids, txt_ids, g, dt)
        b = [m3 * x for x in b]
        v = O.transformer_forward(tr_sd, TR, latents, pe, po, timestep / 1000, img_ids, txt_ids, g, b, None, dt)
        return O.euler_step(v, sig[j], sig[j + 1], latents)

    with torch.no_grad():
        x = lat
        x = step(0, x)
        first = x
        x = step(1, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x = lat
        for i in range(steps):
            x = step(i, x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    d = (first.float() - ours_first.float()).norm() / ours_first.float().norm()
    return dict(value=1000.0 / ms, unit=UNIT, ms_per_step=ms, steps=steps, dtype="bf16",
                kind="oracle/flux_oracle.py (the reference path restated) run by stock torch 2.11: F.linear -> cuBLASLt, "
                     "F.scaled_dot_product_attention -> cuDNN / flash, ATen LayerNorm / RMSNorm / elementwise; same "
                     "weights (shared device tensors), same inputs, device-timed after 2 warm-up steps; runs all 6 "
                     "ControlNet blocks like the reference",
                latents_rel_l2_vs_b200_kernels_step0=float(d))


def elementwise_leg(dev, pk):
    """The HBM-class kernels of the path (north_star: 'achieved HBM GB/s for the elementwise kernels'; SURVEY.md 8d:
    report at the batched cfg-3 size too): Euler, CFG + Euler, mask * scale (+ add), glyph blend at B = 1 and B = 8
    samples of 1024^2.  Every call is timed alone between CUDA events after a 512 MB write that evicts L2."""
    from reptext_b200 import ops
    dt = torch.bfloat16
    flush = torch.empty(128 << 20, dtype=torch.float32, device=dev)
    out = {}

    def t(fn, nbytes, reps=10):
        fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        us = tot / reps * 1000.0
        gbs = nbytes / (us * 1e-6) / 1e9
        return dict(us=round(us, 2), bytes=nbytes, gbs=round(gbs, 1), frac_of_hbm=round(gbs / pk["hbm"], 3))

    for B in (1, 8):
        N, C, D = 4096, 64, 3072
        v, x = torch.randn(B, N, C, device=dev, dtype=dt), torch.randn(B, N, C, device=dev, dtype=dt)
        v2 = torch.randn(2 * B, N, C, device=dev, dtype=dt)
        s6 = torch.randn(B, N, D, device=dev, dtype=dt)
        acc = torch.randn(B, N, D, device=dev, dtype=dt)
        m = torch.rand(N, device=dev).to(dt)
        nz, gl = torch.randn(B, 16, 128, 128, device=dev, dtype=dt), torch.randn(B, 16, 128, 128, device=dev, dtype=dt)
        gm = (torch.rand(B, 16, 128, 128, device=dev) > 0.5).to(torch.uint8)
        e = 2
        T = 512
        xs = torch.randn(B, T + N, D, device=dev, dtype=dt)
        xo = torch.empty_like(xs)
        md = torch.randn(B, 4 * D, device=dev) * 0.3
        ln_groups = [(0, T, md[:, :D], md[:, D:2 * D]), (T, T + N, md[:, 2 * D:3 * D], md[:, 3 * D:])]
        out[f"B{B}"] = dict(
            layernorm_modulate=t(lambda: ops.layernorm_modulate(xs, ln_groups, out=xo), 2 * xs.numel() * e),
            euler_step=t(lambda: ops.euler_step(v, x, 0.9, 0.85), 3 * v.numel() * e),
            cfg_euler_step=t(lambda: ops.cfg_euler_step(v2, x, 3.5, False, 0.9, 0.85), (v2.numel() + 2 * x.numel()) * e),
            mask_scale_add=t(lambda: ops.mask_scale_add(s6, m, acc, 0.9), 3 * s6.numel() * e),
            glyph_init_blend=t(lambda: ops.glyph_init_blend(nz, gl, gm, 0.10, 1.00), 3 * nz.numel() * e + gm.numel()))
    out["note"] = ("layernorm_modulate: the step's LN + AdaLN kernel alone at the joint sequence (512 + 4096 rows), L2 "
                   "flushed before every launch - the `breakdown` figure of the same kernel carries the CUDA-event brackets "
                   "of the profiling pass, and so does this one (an event pair around ONE launch adds ~5 us: ncu times the "
                   "B = 1 launch at 15.2 us, here ~20; at B = 8 the same kernel streams 453 MB at 0.77 of the copy bandwidth); "
                   "bytes = algorithmic reads + writes; at B = 1 the Euler / CFG tensors are 0.5 MB each: the figure is launch "
                   "latency (~2-3 us), not bandwidth; mask_scale_add is the UNFUSED form of the regional-mask multiply and "
                   "multi-line sum (the product fuses both into the zero-linear GEMM epilogue)")
    return out


def make_step(tr, cn, ops, pe, po, cond, mask, img_ids, txt_ids, guidance, sig, tsd, sp_kw):
    dt = torch.bfloat16

    # `cond` / `mask`: one tensor each, or a list with one entry per text line (cfg 2b: the ControlNet runs once per line
    # and the masked samples are summed - pipeline_flux_controlnet.py:1060-1087 - here inside the zero-linear epilogue)
    conds, masks = (cond, mask) if isinstance(cond, (list, tuple)) else ([cond], [mask])

    def one_step(i, latents):
        j = i % STEPS_PER_IMAGE
        kw = dict(hidden_states=latents, encoder_hidden_states=pe, pooled_projections=po,
                  timestep=(tsd[j].expand(1).to(dt)) / 1000, guidance=guidance, img_ids=img_ids, txt_ids=txt_ids)
        acc = None
        for c, m in zip(conds, masks):
            bl, sl = cn(controlnet_cond=c, conditioning_scale=1.0, regional_mask=m, accumulate_into=acc,
                        return_dict=False, **kw, **sp_kw)
            acc = (bl[0]._rt_stacked if bl is not None else None, sl[0]._rt_stacked if sl is not None else None)
        v = tr(controlnet_block_samples=bl, controlnet_single_block_samples=sl, return_dict=False, **kw, **sp_kw)[0]
        return ops.euler_step(v, latents, sig[j], sig[j + 1])
    return one_step


def sp_leg(args, tr, cn, TR, CN, pipe, dev, rank, world, pk, barrier):
    """BASELINE.json configs[4] inside the default multi-GPU run: ONE 1536x1536 sample (N = 9216 image + 512 text tokens),
    tokens and attention heads sharded over all `world` ranks (strong scaling).  Measures, with the SAME weights and inputs:
    the one-GPU step (every rank runs the whole sample itself, max over ranks), the sequence-parallel step, the latents of 2
    free-running steps against the one-GPU result, and the flag barrier alone."""
    import numpy as np
    import torch.distributed as dist
    from reptext_b200 import _lib, ops, parallel
    from reptext_b200._pipeline_common import calculate_shift
    from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import box_mask
    wl = workload("cfg5")
    dt = torch.bfloat16
    H, W, T = wl["H"], wl["W"], wl["T"]
    N = (H // 16) * (W // 16)
    g = torch.Generator().manual_seed(1000)
    lat = torch.randn(1, N, TR["in_channels"], generator=g).to(dt).to(dev)
    pe = torch.randn(1, T, TR["joint_attention_dim"], generator=g).to(dt).to(dev)
    po = torch.randn(1, TR["pooled_projection_dim"], generator=g).to(dt).to(dev)
    cond = torch.randn(1, N, CN["in_channels"] + CN["extra_condition_channels"], generator=g).to(dt).to(dev)
    mask = pipe._regional_masks([box_mask(H, W, (H // 3, H // 3 + H // 6, W // 5, W - W // 5))], dev, dt)[0]
    img_ids = pipe._prepare_latent_image_ids(1, 2 * (H // 16), 2 * (W // 16), dev, dt)
    txt_ids = torch.zeros(T, 3, device=dev, dtype=dt)
    guidance = torch.tensor([3.5], device=dev)
    sch = FlowMatchEulerDiscreteScheduler()
    sc = sch.config
    sch.set_timesteps(sigmas=np.linspace(1.0, 1 / STEPS_PER_IMAGE, STEPS_PER_IMAGE), device=dev,
                      mu=calculate_shift(N, sc.base_image_seq_len, sc.max_image_seq_len, sc.base_shift, sc.max_shift))
    sig, tsd = sch.sigmas.tolist(), sch.timesteps
    K = max(2, min(args.steps, 8))

    def timed(step, x0, prof=False):
        x = x0
        for i in range(3):
            x = step(i, x)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if prof:
            _lib.set_option("profile", 1)
            _lib.profile_reset()
        barrier()
        e0.record()
        x = x0
        for i in range(K):
            x = step(i, x)
            if i == 1:
                two = x
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pr = None
        if prof:
            torch.cuda.synchronize()
            pr = _lib.profile_read()
            _lib.set_option("profile", 0)
            _lib.profile_reset()
        return float(t.item()) / K, two, pr

    one = make_step(tr, cn, ops, pe, po, cond, mask, img_ids, txt_ids, guidance, sig, tsd, {})
    ms1, two1, _ = timed(one, lat)
    sp = parallel.SequenceParallelGroup()
    sh = lambda t, d=1: parallel.shard_tokens(t, rank, world, d)
    stepN = make_step(tr, cn, ops, sh(pe), po, sh(cond), sh(mask.reshape(1, -1, 1)), sh(img_ids, 0), sh(txt_ids, 0),
                      guidance, sig, tsd, dict(sp=sp))
    msN, twoN, _ = timed(stepN, sh(lat))
    sp.check()
    # same-box A/B of the two sequence-parallel design choices (same process, same weights, same inputs): the phase
    # synchronisation as stand-alone barrier kernels instead of inside the neighbouring kernels, and 256-wide GEMM tiles
    # everywhere instead of the run-time tile width; both leave the latents bit-identical
    ab = {}
    for name, opt, val in (("phase_sync_as_barrier_kernels", "sp_sync_kernels", 1), ("static_256_wide_gemm_tiles", "gemm_dyn_bn", -1)):
        _lib.set_option(opt, val)
        ms_ab, two_ab, _ = timed(stepN, sh(lat))
        _lib.set_option(opt, 0)
        sp.check()
        ab[name] = dict(ms_per_step=ms_ab, same_latents=bool(torch.equal(two_ab, twoN)))
    # ... and the pipelines' default on top: the AdaLN vectors of all steps from one pass before the loop
    # (models.build_modulation_table; the figure above recomputes them every step, like the reference)
    ts_all = torch.stack([(tsd[j].expand(1).to(dt)) / 1000 for j in range(STEPS_PER_IMAGE)])
    for net in (cn, tr):
        net.build_modulation_table(ts_all, guidance, po)

    def step_table(i, x):
        for net in (cn, tr):
            net.select_modulation(i % STEPS_PER_IMAGE)
        return stepN(i, x)
    ms_ab, two_ab, _ = timed(step_table, sh(lat))
    for net in (cn, tr):
        net.select_modulation(None)
    sp.check()
    ab["with_modulation_table"] = dict(ms_per_step=ms_ab, same_latents=bool(torch.equal(two_ab, twoN)))
    _, _, prof = timed(stepN, sh(lat), prof=True)      # separate pass: events around every launch
    sp.check()
    twoN = parallel.gather_tokens(twoN)
    err = float((twoN.float() - two1.float()).norm() / two1.float().norm())
    # the flag barrier alone
    for _ in range(10):
        sp.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        sp.barrier()
    e1.record()
    torch.cuda.synchronize()
    bar_us = e0.elapsed_time(e1) / 200 * 1000.0
    sp.close()
    flops = step_flops(TR, CN, N, T, cn_live_layers=cn_live(TR, CN))
    n_sync = 2 * (cn_live(TR, CN) + TR["num_layers"] + TR["num_single_layers"])  # two phase hand-offs per block
    n_bar = 4                                                                    # AdaLN row-shard exchange, 2 per forward
    breakdown = {}
    for k, v in (prof or {}).items():
        tensor = "gemm" in k or "attention" in k
        ach = (v[1] / (v[0] / 1000.0) / (1e12 if tensor else 1e9)) if v[0] else None
        breakdown[k] = dict(ms_per_step=v[0] / K, launches_per_step=v[2] / K, achieved=ach,
                            unit="TFLOP/s" if tensor else "GB/s")
    return dict(workload=wl["name"], n_gpus=world, steps=K, ms_per_step=msN, steps_per_s=1000.0 / msN,
                one_gpu_ms_per_step=ms1, speedup=ms1 / msN, strong_scaling_efficiency=ms1 / msN / world,
                target_ms_per_step=27.1 if world == 8 else None, flops_per_step=flops,
                tensor_util_per_gpu=flops / world / (msN / 1000.0) / 1e12 / pk["bf16_sustained"],
                latents_rel_l2_vs_one_gpu_after_2_steps=err, in_kernel_phase_syncs_per_step=n_sync,
                barrier_kernels_per_step=n_bar, barrier_kernel_us=bar_us, ab=ab, breakdown=breakdown,
                how="same weights, same inputs, same process group as the headline run; one-GPU figure = every rank runs "
                    "the whole sample itself (max over ranks); device-timed, 3 warm-up steps, barrier + synchronize on "
                    "both sides")


def cn_live(TR, CN):
    from reptext_b200.models import FluxControlNetModel
    if CN["num_single_layers"]:
        return CN["num_layers"]
    return FluxControlNetModel.consumed_samples(CN["num_layers"], TR["num_layers"])

# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------
def run_b200(args, wl):
    import torch.distributed as dist
    from reptext_b200 import _lib, models, ops
    from reptext_b200.pipeline_flux_controlnet import FluxControlNetPipeline
    from reptext_b200.pipeline_utils import SyntheticTokenizer
    from reptext_b200 import text_encoders as TE
    from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler
    from reptext_b200.vae import AutoencoderKL
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import box_mask

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a GPU (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    # (A/B aid: RT_OPTIONS=attn_variant=20,... is applied by _lib.lib() at load time)
    dt = torch.bfloat16
    TR, CN, H, W, T = wl["TR"], wl["CN"], wl["H"], wl["W"], wl["T"]
    N = (H // 16) * (W // 16)
    tr = models.FluxTransformer2DModel.random_init(TR, seed=100, dtype=dt, device=dev)
    cn = models.FluxControlNetModel.random_init(CN, seed=101, dtype=dt, device=dev)

    # ---- synthetic inputs, one independent sample (prompt, seed) per rank, staged in PINNED host memory
    g = torch.Generator().manual_seed(1000 + rank)
    pin = lambda t: t.pin_memory()
    h_lat = pin(torch.randn(1, N, TR["in_channels"], generator=g).to(dt))
    h_pe = pin(torch.randn(1, T, TR["joint_attention_dim"], generator=g).to(dt))
    h_po = pin(torch.randn(1, TR["pooled_projection_dim"], generator=g).to(dt))
    h_cond = pin(torch.randn(1, N, CN["in_channels"] + CN["extra_condition_channels"], generator=g).to(dt))
    mask_img = box_mask(H, W, (H // 3, H // 3 + H // 6, W // 5, W - W // 5))
    sch = FlowMatchEulerDiscreteScheduler()
    vae = AutoencoderKL.random_init(seed=102, dtype=dt, device=dev)   # FLUX.1-dev VAE architecture, random weights
    # prompt encoders of the named architectures (T5-v1.1-XXL, CLIP ViT-L/14), random weights drawn on the device; the
    # tokenizers are synthetic (hashing) - their vocabularies are checkpoint files
    if TR["joint_attention_dim"] == TE.T5_XXL_CONFIG["d_model"] and TR["pooled_projection_dim"] == TE.CLIP_L_CONFIG["hidden_size"]:
        t5 = TE.T5EncoderModel(None, TE.random_weights(TE.t5_param_shapes(TE.T5_XXL_CONFIG), 103, dev), dtype=dt, device=dev)
        clip = TE.CLIPTextModel(None, TE.random_weights(TE.clip_param_shapes(TE.CLIP_L_CONFIG), 104, dev), dtype=dt, device=dev)
        tok, tok2 = SyntheticTokenizer("clip", TE.CLIP_L_CONFIG["vocab_size"], 77), SyntheticTokenizer("t5", TE.T5_XXL_CONFIG["vocab_size"], 512)
    else:
        raise SystemExit("bench workloads use the FLUX.1-dev text widths (4096 / 768)")
    pipe = FluxControlNetPipeline(sch, vae, clip, tok, t5, tok2, tr, cn)
    if args.no_mod_table:
        pipe.precompute_modulation = False
    pipe.skip_unconsumed_controlnet_blocks = not args.full_controlnet
    mask = pipe._regional_masks([mask_img], dev, dt)[0]
    lat, pe, po, cond = [t.to(dev, non_blocking=True) for t in (h_lat, h_pe, h_po, h_cond)]
    img_ids = pipe._prepare_latent_image_ids(1, 2 * (H // 16), 2 * (W // 16), dev, dt)
    txt_ids = torch.zeros(T, 3, device=dev, dtype=dt)
    sp = None
    sp_kw = {}
    if wl.get("sp") and world > 1:
        # one sample for the whole job: every rank holds the SAME inputs (seeded alike) and keeps its token shard
        from reptext_b200 import parallel
        g = torch.Generator().manual_seed(1000)
        full = [torch.randn(1, N, TR["in_channels"], generator=g).to(dt), torch.randn(1, T, TR["joint_attention_dim"], generator=g).to(dt),
                torch.randn(1, TR["pooled_projection_dim"], generator=g).to(dt),
                torch.randn(1, N, CN["in_channels"] + CN["extra_condition_channels"], generator=g).to(dt)]
        h_lat, h_pe, h_po, h_cond = [pin(t) for t in full]
        sp = parallel.SequenceParallelGroup()
        sp_kw = dict(sp=sp)
        sh = lambda t, d=1: parallel.shard_tokens(t.to(dev), rank, world, d)
        lat, pe, po, cond = sh(h_lat), sh(h_pe), h_po.to(dev), sh(h_cond)
        mask, img_ids, txt_ids = sh(mask.reshape(1, -1, 1)), sh(img_ids, 0), sh(txt_ids, 0)
    guidance = torch.tensor([3.5], device=dev)
    import numpy as np
    from reptext_b200._pipeline_common import calculate_shift
    sc = sch.config
    sch.set_timesteps(sigmas=np.linspace(1.0, 1 / STEPS_PER_IMAGE, STEPS_PER_IMAGE), device=dev,
                      mu=calculate_shift(N, sc.base_image_seq_len, sc.max_image_seq_len, sc.base_shift, sc.max_shift))
    sig = sch.sigmas.tolist()
    tsd = sch.timesteps

    if not args.full_controlnet:
        # what the pipelines do by default: ControlNet blocks whose sample the transformer never reads are not run
        cn.set_consumer(TR["num_layers"], TR["num_single_layers"])
    L_lines = max(1, args.lines)
    mask_imgs, conds_l, masks_l = [mask_img], [cond], [mask]
    for li in range(1, L_lines):     # cfg 2b: further text lines, each with its own condition and box (stacked below line 1)
        top = (H // 3 + li * (H // 5)) % (H - H // 6)
        mask_imgs.append(box_mask(H, W, (top, top + H // 6, W // 5, W - W // 5)))
        conds_l.append(torch.randn(cond.shape, generator=g).to(dt).to(dev))
        masks_l.append(pipe._regional_masks([mask_imgs[-1]], dev, dt)[0])
    if L_lines > 1:
        assert sp is None, "--lines > 1 is a single-GPU-per-sample measurement"
    one_step = make_step(tr, cn, ops, pe, po, conds_l if L_lines > 1 else cond, masks_l if L_lines > 1 else mask,
                         img_ids, txt_ids, guidance, sig, tsd, sp_kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then EXACTLY K timed steps between barriers, CUDA events on the launch stream
    x = lat
    # W warm-up steps, at least 15 (one second of the cfg-2 step): the power controller needs about that long to settle
    # after the low-power set-up phase - a timed region that starts 0.2 s into the first GEMM-heavy work catches its
    # overshoot (SM clock 1365 MHz against the steady 1440-1500, 70.5 against 68.6 ms per step on one box).  A fixed
    # count: the sequence-parallel ranks must run the same number of steps.
    n_warm = max(args.warmup, 15)
    for i in range(n_warm):
        x = one_step(i, x)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    x = lat
    for i in range(args.steps):
        x = one_step(i, x)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    clocks = sampler.stop() if sampler else None
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    n_samples = 1 if sp is not None else world        # sequence-parallel: the whole job is ONE sample
    value = n_samples * args.steps / (ms_total / 1000.0)
    if sp is not None:
        sp.check()
    finite = bool(torch.isfinite(x.float()).all().item())

    # ---- the same K steps with the pipelines' default on: the AdaLN vectors of all 28 timesteps from ONE pass (built
    #      outside this timed region: informative only, not `value`; the e2e leg below pays for the build)
    with_table_ms = None
    if hasattr(tr, "build_modulation_table"):
        ts_all = torch.stack([(tsd[j].expand(1).to(dt)) / 1000 for j in range(STEPS_PER_IMAGE)])
        for net in (cn, tr):
            net.build_modulation_table(ts_all, guidance, po)

        def step_table(i, x):
            for net in (cn, tr):
                net.select_modulation(i % STEPS_PER_IMAGE)
            return one_step(i, x)
        x2 = lat
        for i in range(3):
            x2 = step_table(i, x2)
        barrier()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        x2 = lat
        for i in range(args.steps):
            x2 = step_table(i, x2)
        t1e.record()
        barrier()
        for net in (cn, tr):
            net.select_modulation(None)
        tt = torch.tensor([t0e.elapsed_time(t1e)], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        with_table_ms = dict(ms_per_step=float(tt.item()) / args.steps, same_latents=bool(torch.equal(x2, x)))
        if sp is not None:
            sp.check()

    # ---- per-class device time of the same K steps (separate pass: events around every launch)
    _lib.set_option("profile", 1)
    _lib.profile_reset()
    x = lat
    for i in range(args.steps):
        x = one_step(i, x)
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.set_option("profile", 0)
    _lib.profile_reset()

    # ---- end to end through the public pipeline API: host (pinned) inputs, one 28-step image, latents read back
    #      to the host after EVERY step (callback_on_step_end, the reference's own per-step tap) and at the end.
    h_tap = torch.empty(1, N // (world if sp is not None else 1), TR["in_channels"], dtype=dt).pin_memory()
    canny = pin(torch.rand(1, 3, H, W, generator=g) * 2 - 1)
    pos = pin((torch.from_numpy(mask_img)[None, None].float() / 255.0) * 2 - 1)
    poss = [pos] + [pin((torch.from_numpy(m)[None, None].float() / 255.0) * 2 - 1) for m in mask_imgs[1:]]
    d2h = [0]

    def tap(p, i, t, kw):
        h_tap.copy_(kw["latents"], non_blocking=True)
        d2h[0] += h_tap.numel() * h_tap.element_size()
        return {}

    def one_image(steps):
        out = pipe(prompt=PROMPT, max_sequence_length=T, height=H, width=W, num_inference_steps=steps,
                   guidance_scale=3.5, control_image=[canny] * L_lines, control_position=poss, control_mask=mask_imgs,
                   controlnet_conditioning_scale=1.0, latents=h_lat, output_type="pt", callback_on_step_end=tap)
        res = out.images.to("cpu", non_blocking=False)     # the decoded image [1, 3, H, W]
        return res

    e2e_value = h2d_step = d2h_step = None
    if not args.no_e2e:
        pipe.enable_sequence_parallel(sp)
        one_image(2)
        barrier()
        d2h[0] = 0
        t0 = time.perf_counter()
        e0.record()
        res = one_image(STEPS_PER_IMAGE)
        e1.record()
        barrier()
        e2e_ms = torch.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0)], device=dev)
        if world > 1:
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        e2e_value = n_samples * STEPS_PER_IMAGE / (float(e2e_ms.item()) / 1000.0)
        h2d_total = (sum(t.numel() * t.element_size() for t in (h_lat, *([canny] * L_lines), *poss))
                     + sum(m.size for m in mask_imgs) * 4 + (T + 77) * 8)
        h2d_step = h2d_total / STEPS_PER_IMAGE
        d2h_step = (d2h[0] + res.numel() * res.element_size()) / STEPS_PER_IMAGE

    # ---- the prompt encoders on their own: once per image, device-timed
    text_info = None
    if rank == 0 and not args.no_e2e:
        ids5 = tok2([PROMPT], padding="max_length", max_length=T, truncation=True).input_ids.to(dev)
        idsc = tok([PROMPT], padding="max_length", max_length=77, truncation=True).input_ids.to(dev)
        def _tt(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        c5 = TE.T5_XXL_CONFIG
        t5_flop = c5["num_layers"] * (2.0 * T * (4 * c5["d_model"] * c5["d_model"] + 3 * c5["d_model"] * c5["d_ff"])
                                      + 4.0 * T * T * c5["d_model"])
        t5_ms, clip_ms = _tt(lambda: t5(ids5)), _tt(lambda: clip(idsc))
        text_info = dict(t5_xxl_ms=t5_ms, t5_xxl_tflops=t5_flop / t5_ms / 1e9, clip_l_ms=clip_ms, tokens=T,
                         note="T5EncoderModel / CLIPTextModel drop-ins (reptext_b200/text_encoders.py), random weights, "
                              "synthetic tokenizer; inside the e2e figure once per image")

    # ---- the VAE on its own: two encodes per text line + one decode per image, device-timed
    vae_info = None
    if rank == 0 and not args.no_e2e:
        zt = torch.randn(1, 16, H // 8, W // 8, device=dev, dtype=dt)
        it = canny.to(dev, dt)
        def _t(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        nl = _lib.launch_count()
        vae.decode(zt)
        nl = _lib.launch_count() - nl
        vae_info = dict(encode_ms=_t(lambda: vae.encode(it)), decode_ms=_t(lambda: vae.decode(zt)), decode_launches=nl,
                        note="AutoencoderKL drop-in (reptext_b200/vae.py), FLUX.1-dev VAE architecture, random weights; "
                             "inside the e2e figure: 2 encodes + 1 decode per image")

    # ---- the only collective: gather the output latents of all ranks (after the timed regions)
    if sp is not None:
        from reptext_b200 import parallel
        x = parallel.gather_tokens(x)
        sp.close()
    elif world > 1:
        outs = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(outs, x)

    pk = peaks()
    # ---- the kernel to beat, same box, same inputs: the reference path through stock torch in bf16
    gpu_base = None
    if world == 1 and sp is None and not args.no_gpu_baseline and L_lines == 1:
        try:
            gpu_base = gpu_baseline_leg(tr, cn, TR, CN, lat, pe, po, cond, mask, img_ids, txt_ids, sch.sigmas, tsd,
                                        max(2, min(args.steps, 6)), one_step(0, lat))
        except Exception as e:  # a baseline that cannot run must not take the product's line with it
            gpu_base = dict(unavailable=repr(e)[:300])
        torch.cuda.empty_cache()
    elementwise = elementwise_leg(dev, pk) if (rank == 0 and not args.no_e2e) else None
    # ---- BASELINE.json configs[4] next to the headline: one 1536^2 sample sequence-parallel over all ranks
    sp_rec = None
    if world > 1 and sp is None and not args.no_sp and args.workload == "cfg2":
        try:
            sp_rec = sp_leg(args, tr, cn, TR, CN, pipe, dev, rank, world, pk, barrier)
        except Exception as e:
            sp_rec = dict(unavailable=repr(e)[:300])

    if rank == 0:
        live = None if args.full_controlnet else cn_live(TR, CN)
        flops = step_flops(TR, CN, N, T, lines=L_lines, cn_live_layers=live)
        flops_reference = step_flops(TR, CN, N, T, lines=L_lines)
        gpus_per_sample = world if sp is not None else 1
        gem = prof.get("gemm_tcgen05", (0.0, 0.0, 0))
        roof = None
        # ncu dram__bytes_read + write per GEMM launch, averaged over the GEMM launches of one cfg2 step (tools/gpu.sh traffic)
        traffic, traffic_file = None, None
        if args.workload == "cfg2" and L_lines == 1:
            for name in ("r2_gemm_dram_traffic.json", "r1_gemm_dram_traffic.json"):
                tpath = os.path.join(ROOT, "profiles", name)
                if not os.path.exists(tpath):
                    continue
                tj = json.load(open(tpath))
                fam = tj.get("families", {}).get("gemm_tc_kernel") or tj.get("banded_long_k") or tj.get("no_l2_hints")
                if fam:
                    traffic, traffic_file = fam["bytes_per_launch"], name
                    break
        if gem[0] > 0:
            ach = gem[1] / (gem[0] / 1000.0) / 1e12
            roof = dict(bound="tensor", kernel="gemm_tc_kernel (tcgen05 + TMA, fused epilogues)", achieved=ach,
                        peak=pk["bf16_sustained"], unit="TFLOP/s", frac=ach / pk["bf16_sustained"], traffic=traffic,
                        traffic_unit=(f"bytes of DRAM traffic per launch (ncu, profiles/{traffic_file}); "
                                      "algorithmic operand bytes per launch: 203e6") if traffic else None,
                        peak_source=pk["source"] + ", sustained figure (kernel timed inside a long step)",
                        frac_of_burst=ach / pk["bf16_burst"], launches=gem[2], ms_per_step=gem[0] / args.steps,
                        flops_per_step=gem[1] / args.steps)
        def _cls(k, v):
            tensor = "gemm" in k or "attention" in k
            ach = (v[1] / (v[0] / 1000.0) / (1e12 if tensor else 1e9)) if v[0] else None
            peak = pk["bf16_sustained"] if tensor else pk["hbm"]
            return dict(ms_per_step=v[0] / args.steps, launches_per_step=v[2] / args.steps, achieved=ach,
                        unit="TFLOP/s" if tensor else "GB/s", bound="tensor" if tensor else "hbm", peak=peak,
                        frac=(ach / peak) if ach else None)
        breakdown = {k: _cls(k, v) for k, v in prof.items()}   # every kernel class against its own roofline
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            v_cpu, sample = cpu_oracle_steps_per_s(wl)
            cpu = dict(value=v_cpu, unit=UNIT, cores=torch.get_num_threads(), kind="port", sample=sample)
        step_ms = ms_total / args.steps
        line = dict(metric=METRIC.replace("1024x1024", f"{H}x{W}"), value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=n_warm,
                    ms_per_step=step_ms, higher_is_better=True, scaling="strong" if sp is not None else "weak",
                    vs_baseline=None, dtype="bf16",
                    data="synthetic (random-init weights of the named architecture, randn latents / embeddings)",
                    config=dict(workload=wl["name"] if L_lines == 1 else wl["name"].replace("1 text line", f"{L_lines} text lines (cfg 2b)"),
                                parallelism=(f"sp{world} (one sample; tokens and attention heads sharded; peer stores + flag "
                                             "barriers, NCCL only gathers the final latents)" if sp is not None else
                                             f"dp{world} (independent samples, no data-path collective)"),
                                l2="not flushed: each step streams 32 GB of weights, far larger than the 126 MB L2",
                                flops_per_step=flops, flops_per_step_reference=flops_reference,
                                controlnet_blocks_run=(CN["num_layers"] if live is None else live),
                                controlnet_note=("the pipelines' default: ControlNet blocks whose sample the transformer "
                                                 "never reads (i // ceil(19 / 6): sample 5) are not run; latents are "
                                                 "bit-identical (tests/test_fullsize_gpu.py); --full-controlnet runs all 6; "
                                                 "utilisation figures count only the FLOPs executed"),
                                step_invariant_cache=("OFF in the device-timed loop (`value`: every step recomputes "
                                                      "context_embedder, the rotary table and the guidance / pooled "
                                                      "linears, like the reference); ON in `e2e`, the pipelines' default "
                                                      "(once per image, bit-identical latents, SURVEY 8f.2: 0.15 % of a step)"),
                                modulation_table=("OFF in the device-timed loop (`value`: every step runs the timestep / "
                                                  "guidance / pooled embedders and streams the 6.5 GB of AdaLN weights, "
                                                  "like the reference); ON in `e2e`, the pipelines' default: the AdaLN "
                                                  "vectors of all 28 steps are computed in one pass inside the timed "
                                                  "call, before the loop (weights read once per image, bit-identical "
                                                  "latents, tests/test_pipeline_gpu.py)"),
                                device_loop_with_modulation_table=with_table_ms, warmup_steps_run=n_warm,
                                images_per_s=value / STEPS_PER_IMAGE,
                                tensor_util_whole_step=flops / gpus_per_sample / (step_ms / 1000.0) / 1e12 / pk["bf16_sustained"],
                                finite_output=finite),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d_step, d2h_bytes_per_step=d2h_step,
                             how="FluxControlNetPipeline.__call__(prompt=str, output_type='pt'), one 28-step image from "
                                 "host inputs: tokenise, CLIP-L + T5-XXL prompt encode, VAE encode of the Canny and position "
                                 "images, 28 denoise steps with the latents copied to the host after every step, VAE "
                                 "decode, image copied to the host"),
                    vae=vae_info, text_encoders=text_info,
                    gpu_launches=launches, clocks=clocks, roofline=roof, cpu_baseline=cpu, gpu_baseline=gpu_base,
                    breakdown=breakdown, elementwise=elementwise, sp=sp_rec)
        if gpu_base and gpu_base.get("value"):
            line["gpu_baseline"]["b200_kernels_over_stock_torch"] = value / gpu_base["value"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the pipeline end-to-end leg (profiling runs only)")
    ap.add_argument("--full-controlnet", action="store_true", help="run every ControlNet block like the reference (A/B of the unconsumed-block skip)")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-mod-table", action="store_true", help="e2e leg: AdaLN vectors computed per step instead of once per image (A/B of pipe.precompute_modulation)")
    ap.add_argument("--lines", type=int, default=1, help="text lines per image (cfg 2b = 2: the ControlNet runs once per line); the baselines are measured at 1 only")
    ap.add_argument("--no-sp", action="store_true", help="skip the cfg5 sequence-parallel record of multi-GPU runs")
    args = ap.parse_args()
    wl = workload(args.workload)
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
