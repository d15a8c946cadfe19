cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 examples/sweep.py --config small --samples 9 --steps 3 2>&1 | grep -v "^\*\|OMP" | tail -4
