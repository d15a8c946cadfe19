#!/usr/bin/env python
"""What the two per-block exchanges of the sequence-parallel mode would cost as NCCL collectives (the baseline the
fused peer-store epilogues replace): all_to_all_single of the q|k|v shard and of the attention-output shard at the cfg-5
sizes, timed with CUDA events, max over ranks.  Run with torchrun on N GPUs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S, D, blocks = 9728, 3072, 63
    S_loc = S // world
    qkv_in = torch.randn(world, S_loc, 3 * D // world, device="cuda").to(torch.bfloat16)
    qkv_out = torch.empty_like(qkv_in)
    o_in = torch.randn(world, S_loc, D // world, device="cuda").to(torch.bfloat16)
    o_out = torch.empty_like(o_in)

    def timeit(fn, iters=50):
        for _ in range(5): fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_qkv = timeit(lambda: dist.all_to_all_single(qkv_out, qkv_in))
    t_o = timeit(lambda: dist.all_to_all_single(o_out, o_in))
    t_bar = timeit(lambda: dist.barrier())
    from reptext_b200 import parallel
    sp = parallel.SequenceParallelGroup()
    t_flag = timeit(lambda: sp.barrier(), iters=200)
    sp.check()
    if rank == 0:
        mb = lambda t: t.numel() * 2 / 1e6
        print(f"world {world}: all_to_all q|k|v ({mb(qkv_in):.1f} MB per rank) {t_qkv*1e3:.0f} us | attention output "
              f"({mb(o_in):.1f} MB per rank) {t_o*1e3:.0f} us | per block {(t_qkv+t_o)*1e3:.0f} us | per step "
              f"({blocks} blocks) {(t_qkv+t_o)*blocks:.2f} ms  (+ pack / unpack passes; NCCL barrier {t_bar*1e3:.0f} us; "
              f"the flag barrier of rt_sp_barrier {t_flag*1e3:.1f} us, 126 + 4 per step)")
    sp.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
