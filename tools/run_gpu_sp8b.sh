# cfg5 (one 1536x1536 sample, sequence-parallel over all GPUs of the box) after the GEMM change
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -W ignore -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload cfg5 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/sp${N}_cfg5_v2.json 2> gpurun_out/sp${N}_cfg5_v2.err; echo "cfg5 sp$N rc=$?"; grep '^{' gpurun_out/sp${N}_cfg5_v2.json | cut -c1-330
