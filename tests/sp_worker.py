"""Worker of tests/test_sp_gpu.py::test_multi_process_sequence_parallel_pipeline_matches_single_gpu (run by torchrun,
one process per GPU).  Every rank builds the same small model pair and runs the same T2I pipeline call twice:
single-GPU, then sequence-parallel over all ranks; the latents must agree."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from reptext_b200 import config, models, parallel
    from reptext_b200.pipeline_flux_controlnet import FluxControlNetPipeline
    from reptext_b200.pipeline_utils import SyntheticTextEncoders, SyntheticVAE
    from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler
    from util import box_mask, rel_l2

    rank, world, local = parallel.init_from_env("nccl")
    dev = torch.device("cuda", local)
    dt = torch.bfloat16
    TR, CN = config.SP8_TRANSFORMER, config.SP8_CONTROLNET
    tr = models.FluxTransformer2DModel.random_init(TR, seed=100, dtype=dt, device=dev)
    cn = models.FluxControlNetModel.random_init(CN, seed=101, dtype=dt, device=dev)
    # 1024 image + 512 text tokens: at world 2 / 4 / 8 every rank's text rows are a multiple of 64 and its image rows a
    # multiple of 128 - the shapes for which the sharded run must equal the single-GPU run bit for bit
    H, W, T = 512, 512, 512
    N = (H // 16) * (W // 16)
    g = torch.Generator().manual_seed(5)
    pe = torch.randn(1, T, TR["joint_attention_dim"], generator=g).to(dt)
    po = torch.randn(1, TR["pooled_projection_dim"], generator=g).to(dt)
    lat = torch.randn(1, N, TR["in_channels"], generator=g).to(dt)
    canny = torch.rand(1, 3, H, W, generator=g) * 2 - 1
    mask_img = box_mask(H, W, (H // 3, H // 3 + H // 6, W // 5, W - W // 5))
    pos = (torch.from_numpy(mask_img)[None, None].float() / 255.0) * 2 - 1
    pipe = FluxControlNetPipeline(FlowMatchEulerDiscreteScheduler(), SyntheticVAE(dtype=dt, device=dev),
                                  SyntheticTextEncoders(TR["joint_attention_dim"], TR["pooled_projection_dim"], dt, dev),
                                  None, None, None, tr, cn)

    def run():
        return pipe(prompt_embeds=pe, pooled_prompt_embeds=po, height=H, width=W, num_inference_steps=4,
                    guidance_scale=3.5, control_image=[canny], control_position=[pos], control_mask=[mask_img],
                    controlnet_conditioning_scale=1.0, latents=lat.clone(), output_type="latent").images

    single = run()
    sp = parallel.SequenceParallelGroup()
    for _ in range(3):                      # a few barriers on their own first: epochs, flags, mappings
        sp.barrier()
    sp.check()
    pipe.enable_sequence_parallel(sp)
    multi = run()
    multi2 = run()                          # the epoch counters keep growing across calls
    # the phase barriers as stand-alone kernels instead of at the head of the attention / output-projection kernels
    # (csrc/sp_sync.cuh): same protocol, same epochs - the two forms mix in one stream - same bits
    from reptext_b200 import _lib
    _lib.set_option("sp_sync_kernels", 1)
    multi_k = run()
    _lib.set_option("sp_sync_kernels", 0)
    multi_b = run()
    forms = torch.equal(multi_k, multi) and torch.equal(multi_b, multi)
    print(f"[rank {rank}] barrier forms (kernels / in-kernel / back) agree: {forms}", flush=True)
    # one sequence-parallel denoising step captured into a CUDA graph on every rank and replayed on new latents: the
    # in-kernel barriers take their epochs from the flag block on the device, so a replay synchronises like the eager run
    from reptext_b200 import ops
    shard = lambda t, d=1: parallel.shard_tokens(t.to(dev), rank, world, d)
    ids_i = pipe._prepare_latent_image_ids(1, 2 * (H // 16), 2 * (W // 16), dev, dt)
    ids_t = torch.zeros(T, 3, device=dev, dtype=dt)
    condp = torch.randn(1, N, CN["in_channels"] + CN["extra_condition_channels"], generator=g).to(dt)
    tt, gg = torch.tensor([0.62], device=dev, dtype=dt), torch.tensor([3.5], device=dev)

    pe_s, po_d, ii_s, ti_s, cond_s = (shard(pe).contiguous(), po.to(dev), shard(ids_i, 0).float().contiguous(),
                                      shard(ids_t, 0).float().contiguous(), shard(condp).contiguous())

    def sp_step(z):
        kw = dict(hidden_states=z, encoder_hidden_states=pe_s, pooled_projections=po_d, timestep=tt,
                  guidance=gg.to(dt), img_ids=ii_s, txt_ids=ti_s, sp=sp)
        bl, _ = cn(controlnet_cond=cond_s, conditioning_scale=0.9, return_dict=False, **kw)
        v = tr(controlnet_block_samples=bl, return_dict=False, **kw)[0]
        return ops.euler_step(v, z, 0.62, 0.55)

    z1, z2 = shard(lat).contiguous(), shard(torch.randn(1, N, TR["in_channels"], generator=g).to(dt)).contiguous()
    eager1, eager2 = sp_step(z1).clone(), sp_step(z2).clone()
    static_in = z1.clone()
    cs = torch.cuda.Stream(dev)
    cs.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(cs):
        sp_step(static_in)
    torch.cuda.current_stream(dev).wait_stream(cs)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=cs):
        static_out = sp_step(static_in)
    graph.replay()
    torch.cuda.synchronize()
    g_ok = torch.equal(static_out, eager1)
    static_in.copy_(z2)
    graph.replay()
    torch.cuda.synchronize()
    g_ok = g_ok and torch.equal(static_out, eager2)
    sp.check()
    print(f"[rank {rank}] CUDA-graph replay of a sequence-parallel step equals the eager step: {g_ok}", flush=True)
    del graph
    pipe.enable_sequence_parallel(None)
    # ---- the inpaint pipeline (second ControlNet, true CFG at effective batch 2) the same way
    from reptext_b200.pipeline_flux_controlnet_inpaint import FluxControlNetPipeline as InpaintPipeline
    cni = models.FluxControlNetModel.random_init(config.SP8_INPAINT_CONTROLNET, seed=103, dtype=dt, device=dev)
    ipipe = InpaintPipeline(FlowMatchEulerDiscreteScheduler(), SyntheticVAE(dtype=dt, device=dev),
                            SyntheticTextEncoders(TR["joint_attention_dim"], TR["pooled_projection_dim"], dt, dev),
                            None, None, None, tr, cn, cni)
    npe = torch.randn(1, T, TR["joint_attention_dim"], generator=g).to(dt)
    npo = torch.randn(1, TR["pooled_projection_dim"], generator=g).to(dt)
    src = torch.rand(1, 3, H, W, generator=g) * 2 - 1

    def run_inpaint():
        torch.manual_seed(7)                # the VAE posterior is sampled with the global RNG
        return ipipe(prompt_embeds=pe, pooled_prompt_embeds=po, negative_prompt_embeds=npe,
                     negative_pooled_prompt_embeds=npo, height=H, width=W, num_inference_steps=3, guidance_scale=3.5,
                     true_guidance_scale=3.0, control_image=[canny], control_position=[pos], control_mask=[mask_img],
                     control_glyph=src, control_image_inpaint=src, control_mask_inpaint=mask_img,
                     controlnet_conditioning_scale=1.0, controlnet_conditioning_scale_inpaint=0.8,
                     generator=torch.Generator(device=dev).manual_seed(11), output_type="latent").images

    i_single = run_inpaint()
    ipipe.enable_sequence_parallel(sp)
    i_multi = run_inpaint()
    ipipe.enable_sequence_parallel(None)
    i_err = rel_l2(i_multi, i_single)
    print(f"[rank {rank}] inpaint sp world={world} rel_l2 vs single GPU = {i_err:.3e}", flush=True)
    err = rel_l2(multi, single)
    finite, repeat = torch.isfinite(multi.float()).all().item(), torch.equal(multi, multi2)
    print(f"[rank {rank}] sp world={world} rel_l2 vs single GPU = {err:.3e} finite={finite} repeatable={repeat} "
          f"rel_l2(run2, run1)={rel_l2(multi2, multi):.3e}", flush=True)
    exact = torch.equal(multi, single) and torch.equal(i_multi, i_single)
    print(f"[rank {rank}] sp world={world} bit-identical to the single-GPU run: {exact}", flush=True)
    ok = finite and exact and repeat and forms and g_ok and torch.isfinite(i_multi.float()).all().item()
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    # every rank holds the same gathered latents
    ref = multi.clone()
    dist.broadcast(ref, 0)
    same = torch.equal(ref, multi)
    sp.close()
    if rank == 0:
        print(f"sp world={world} rel_l2 vs single GPU = {err:.3e} identical_across_ranks={same}")
        if flag.item() == 1 and same:
            print("SP_WORKER_OK")
    dist.destroy_process_group()
    if not (flag.item() == 1 and same):
        sys.exit(1)


if __name__ == "__main__":
    main()
