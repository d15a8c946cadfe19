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
    ok = finite and exact and repeat and torch.isfinite(i_multi.float()).all().item()
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
