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


def step_flops(TR, CN, N, T, lines=1):
    """2*M*N*K per GEMM, 4*S^2*D per attention; norms / softmax / elementwise not counted."""
    def model(c, kind):
        D = c["num_attention_heads"] * c["attention_head_dim"]
        S = N + T
        dbl = 24 * S * D * D + 4 * S * S * D
        f = c["num_layers"] * dbl + c["num_single_layers"] * dbl
        f += 2 * N * D * c["in_channels"] + 2 * T * D * c["joint_attention_dim"]
        if kind == "cn":
            f += 2 * N * D * (c["in_channels"] + c["extra_condition_channels"])
            f += (c["num_layers"] + c["num_single_layers"]) * 2 * N * D * D
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
def cpu_oracle_steps_per_s(wl, repeats=1):
    """Times ONE double-stream and ONE single-stream block of the oracle at the full sequence length in fp32 on all
    host threads and extrapolates to the step: (L_tr + L_cn) doubles + L_single singles (+ ControlNet zero-linears).
    A full step is 82.7 TFLOP - minutes on host cores - so the sample is bounded (task statement, section 4)."""
    from oracle import flux_oracle as O
    from reptext_b200 import weights
    TR, CN = wl["TR"], wl["CN"]
    N, T = (wl["H"] // 16) * (wl["W"] // 16), wl["T"]
    D = TR["num_attention_heads"] * TR["attention_head_dim"]
    one = dict(TR, num_layers=1, num_single_layers=1)
    sd = {k: v for k, v in weights.random_state_dict(one, "transformer", seed=0).items()
          if k.startswith(("transformer_blocks.0.", "single_transformer_blocks.0."))}
    g = torch.Generator().manual_seed(0)
    x, c = torch.randn(1, N, D, generator=g), torch.randn(1, T, D, generator=g)
    temb = torch.randn(1, D, generator=g)
    ids = torch.cat([torch.zeros(T, 3), O.prepare_latent_image_ids(2 * (wl["H"] // 16), 2 * (wl["W"] // 16))])
    rope = O.rope_table(ids, TR["axes_dims_rope"])
    Wz = torch.randn(D, D, generator=g) * D ** -0.5
    with torch.no_grad():
        td = ts = tz = 1e30
        for _ in range(repeats):
            t0 = time.perf_counter(); O.double_block(sd, "transformer_blocks.0.", x, c, temb, rope, TR["num_attention_heads"])
            t1 = time.perf_counter(); O.single_block(sd, "single_transformer_blocks.0.", x, c, temb, rope, TR["num_attention_heads"])
            t2 = time.perf_counter(); torch.nn.functional.linear(x, Wz)
            t3 = time.perf_counter()
            td, ts, tz = min(td, t1 - t0), min(ts, t2 - t1), min(tz, t3 - t2)
    n_d = TR["num_layers"] + CN["num_layers"]
    n_s = TR["num_single_layers"] + CN["num_single_layers"]
    step_s = n_d * td + n_s * ts + (CN["num_layers"] + CN["num_single_layers"]) * tz
    sample = (f"oracle (fp32 torch CPU port of the diffusers path) timed on 1 double block ({td:.2f} s) + 1 single block "
              f"({ts:.2f} s) + 1 zero-linear ({tz:.2f} s) at S={N + T}, D={D}; step = {n_d} doubles + {n_s} singles "
              f"+ {CN['num_layers']} zero-linears, extrapolated")
    return 1.0 / step_s, sample


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    vals = []
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_oracle_steps_per_s(wl)
    sample = ""
    t_begin = time.perf_counter()
    for _ in range(max(1, args.steps)):
        v, sample = cpu_oracle_steps_per_s(wl)
        vals.append(v)
        if time.perf_counter() - t_begin > 150:
            break
    v = sum(vals) / len(vals)
    cores = torch.get_num_threads()
    line = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=len(vals), warmup=args.warmup,
                ms_per_step=1000.0 / v, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=dict(workload=wl["name"], note="CPU arm does not scale with --gpus"),
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                images_per_s=v / STEPS_PER_IMAGE)
    print(json.dumps(line), flush=True)


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

    def one_step(i, latents):
        t = tsd[i % STEPS_PER_IMAGE]
        kw = dict(hidden_states=latents, encoder_hidden_states=pe, pooled_projections=po,
                  timestep=(t.expand(1).to(dt)) / 1000, guidance=guidance, img_ids=img_ids, txt_ids=txt_ids)
        bl, sl = cn(controlnet_cond=cond, conditioning_scale=1.0, regional_mask=mask, return_dict=False, **kw, **sp_kw)
        v = tr(controlnet_block_samples=bl, controlnet_single_block_samples=sl, return_dict=False, **kw, **sp_kw)[0]
        j = i % STEPS_PER_IMAGE
        return ops.euler_step(v, latents, sig[j], sig[j + 1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then EXACTLY K timed steps between barriers, CUDA events on the launch stream
    x = lat
    for i in range(max(args.warmup, 3)):
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
    d2h = [0]

    def tap(p, i, t, kw):
        h_tap.copy_(kw["latents"], non_blocking=True)
        d2h[0] += h_tap.numel() * h_tap.element_size()
        return {}

    def one_image(steps):
        out = pipe(prompt=PROMPT, max_sequence_length=T, height=H, width=W, num_inference_steps=steps,
                   guidance_scale=3.5, control_image=[canny], control_position=[pos], control_mask=[mask_img],
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
        h2d_total = sum(t.numel() * t.element_size() for t in (h_lat, canny, pos)) + mask_img.size * 4 + (T + 77) * 8
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

    if rank == 0:
        pk = peaks()
        flops = step_flops(TR, CN, N, T)
        gpus_per_sample = world if sp is not None else 1
        gem = prof.get("gemm_tcgen05", (0.0, 0.0, 0))
        roof = None
        traffic = None   # ncu dram__bytes_read + write per GEMM launch, averaged over the 188 launches of one cfg2 step
        tpath = os.path.join(ROOT, "profiles", "r1_gemm_dram_traffic.json")
        if args.workload == "cfg2" and os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = (tj.get("banded_long_k") or tj["no_l2_hints"])["bytes_per_launch"]
        if gem[0] > 0:
            ach = gem[1] / (gem[0] / 1000.0) / 1e12
            roof = dict(bound="tensor", kernel="gemm_tc_kernel (tcgen05 + TMA, fused epilogues)", achieved=ach,
                        peak=pk["bf16_sustained"], unit="TFLOP/s", frac=ach / pk["bf16_sustained"], traffic=traffic,
                        traffic_unit="bytes of DRAM traffic per launch (ncu, profiles/r1_gemm_dram_traffic.json); "
                                     "algorithmic operand bytes per launch: 203e6",
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
        line = dict(metric=METRIC.replace("1024x1024", f"{H}x{W}"), value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=step_ms, higher_is_better=True, scaling="strong" if sp is not None else "weak",
                    vs_baseline=None, dtype="bf16",
                    data="synthetic (random-init weights of the named architecture, randn latents / embeddings)",
                    config=dict(workload=wl["name"],
                                parallelism=(f"sp{world} (one sample; tokens and attention heads sharded; peer stores + flag "
                                             "barriers, NCCL only gathers the final latents)" if sp is not None else
                                             f"dp{world} (independent samples, no data-path collective)"),
                                l2="not flushed: each step streams 32 GB of weights, far larger than the 126 MB L2",
                                flops_per_step=flops, images_per_s=value / STEPS_PER_IMAGE,
                                tensor_util_whole_step=flops / gpus_per_sample / (step_ms / 1000.0) / 1e12 / pk["bf16_sustained"],
                                finite_output=finite),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d_step, d2h_bytes_per_step=d2h_step,
                             how="FluxControlNetPipeline.__call__(prompt=str, output_type='pt'), one 28-step image from "
                                 "host inputs: tokenise, CLIP-L + T5-XXL prompt encode, VAE encode of the Canny and position "
                                 "images, 28 denoise steps with the latents copied to the host after every step, VAE "
                                 "decode, image copied to the host"),
                    vae=vae_info, text_encoders=text_info,
                    gpu_launches=launches, clocks=clocks, roofline=roof, cpu_baseline=cpu, breakdown=breakdown)
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
    args = ap.parse_args()
    wl = workload(args.workload)
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
