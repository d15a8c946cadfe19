set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12
timeout 600 python -m pytest tests/test_sp_gpu.py -m gpu -x -q -k multi_process 2>&1 | tail -40 > gpurun_out/sp2_test.log; cat gpurun_out/sp2_test.log
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 4 --warmup 3 --workload cfg5 > gpurun_out/bench_cfg5_sp$N.json 2> gpurun_out/bench_cfg5_sp$N.err; tail -c 1500 gpurun_out/bench_cfg5_sp$N.json; tail -5 gpurun_out/bench_cfg5_sp$N.err
timeout 600 python bench.py --steps 4 --warmup 3 --workload cfg5 --no-cpu-baseline > gpurun_out/bench_cfg5_1gpu.json 2> gpurun_out/bench_cfg5_1gpu.err; tail -c 1500 gpurun_out/bench_cfg5_1gpu.json
timeout 300 python tools/sp_scatter_probe.py > gpurun_out/sp_scatter_probe.log 2>&1; cat gpurun_out/sp_scatter_probe.log
