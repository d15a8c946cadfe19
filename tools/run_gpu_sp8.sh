cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -W ignore -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 examples/sweep.py --config flux-dev --samples 64 --steps 28 --height 1024 --width 1024 > gpurun_out/sweep_cfg3_$N.log 2>&1; grep "^{" gpurun_out/sweep_cfg3_$N.log | tail -1
