cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_sp_gpu.py -m gpu -x -q -k multi_process 2>&1 | tail -3; grep "rel_l2" gpurun_out/sp_worker.log | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 --workload cfg5 > gpurun_out/bench_cfg5_sp${N}_final.json 2> gpurun_out/bench_cfg5_sp${N}_final.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_cfg2_dp${N}_final.json 2> gpurun_out/bench_cfg2_dp${N}_final.err
python - <<EOF
import json
for f in ['gpurun_out/bench_cfg5_sp${N}_final.json','gpurun_out/bench_cfg2_dp${N}_final.json']:
    d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    print(f, round(d['ms_per_step'],2), 'ms/step', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'], {k:(round(v['ms_per_step'],2), round(v['achieved'] or 0)) for k,v in d['breakdown'].items()})
EOF
