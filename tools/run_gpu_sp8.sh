cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/nccl_a2a_probe.py > gpurun_out/nccl_a2a_probe_$N.log 2>&1; grep -v "^\*\|OMP" gpurun_out/nccl_a2a_probe_$N.log | tail -5
