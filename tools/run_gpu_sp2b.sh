# 2-GPU sanity after the GEMM change: the multi-process sequence-parallel test, cfg5 on two ranks, cfg2 data-parallel on two ranks
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_sp_gpu.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --workload cfg5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/sp2_cfg5.json 2> gpurun_out/sp2_cfg5.err; echo "cfg5 sp2 rc=$?"; grep '^{' gpurun_out/sp2_cfg5.json | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/dp2_cfg2.json 2> gpurun_out/dp2_cfg2.err; echo "cfg2 dp2 rc=$?"; grep '^{' gpurun_out/dp2_cfg2.json | cut -c1-330
