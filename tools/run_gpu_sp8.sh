cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_sp_gpu.py -m gpu -x -q -k multi_process 2>&1 | tail -5 > gpurun_out/sp${N}_test.log; cat gpurun_out/sp${N}_test.log; grep "rel_l2" gpurun_out/sp_worker.log | head -3
for n in $N 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 6 --warmup 3 --workload cfg5 > gpurun_out/bench_cfg5_sp$n.json 2> gpurun_out/bench_cfg5_sp$n.err
python - <<EOF
import json
d=json.loads([l for l in open('gpurun_out/bench_cfg5_sp$n.json') if l.startswith('{')][-1])
print('cfg5 sp$n', round(d['ms_per_step'],2), 'ms/step', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), d['clocks'], {k:(round(v['ms_per_step'],2), round(v['achieved'] or 0)) for k,v in d['breakdown'].items()})
EOF
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_cfg2_dp$N.json 2> gpurun_out/bench_cfg2_dp$N.err
python - <<EOF
import json
d=json.loads([l for l in open('gpurun_out/bench_cfg2_dp$N.json') if l.startswith('{')][-1])
print('cfg2 dp$N', round(d['ms_per_step'],2), 'ms/step', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), d['clocks'])
EOF
