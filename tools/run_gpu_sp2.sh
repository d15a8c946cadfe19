cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sp_gpu.py -m gpu -x -q 2>&1 | tail -4; grep "rel_l2" gpurun_out/sp_worker.log | head -2
N=$(nvidia-smi -L | wc -l)
for o in none no_pdl=1; do
RT_OPTIONS=$([ $o = none ] && echo "" || echo $o) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 6 --warmup 3 --workload cfg5 > gpurun_out/bench_cfg5_sp${N}_$o.json 2> gpurun_out/bench_cfg5_sp${N}_$o.err
python - <<EOF
import json
d=json.loads([l for l in open('gpurun_out/bench_cfg5_sp${N}_$o.json') if l.startswith('{')][-1])
print('cfg5 sp$N $o', round(d['ms_per_step'],2), 'ms/step', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'])
EOF
done
