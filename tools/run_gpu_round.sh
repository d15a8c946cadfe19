cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/attn_sweep.py 2,10,17,18,19 > gpurun_out/attn_sweep5.log 2>&1; cat gpurun_out/attn_sweep5.log
for v in 0 15 16; do
RT_OPTIONS=attn_variant=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench11_v$v.json 2> gpurun_out/bench11_v$v.err
python - <<EOF
import json
d=json.loads([l for l in open('gpurun_out/bench11_v$v.json') if l.startswith('{')][-1])
print('variant $v', round(d['ms_per_step'],2), {k:(round(v['ms_per_step'],2), round(v['achieved'] or 0)) for k,v in d['breakdown'].items()})
EOF
done
