cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for o in none ln_impl=1 none ln_impl=1; do
RT_OPTIONS=$([ $o = none ] && echo "" || echo $o) timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench15_$o.json 2> gpurun_out/bench15_$o.err
python - <<EOF
import json
d=json.loads([l for l in open('gpurun_out/bench15_$o.json') if l.startswith('{')][-1])
print('$o', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k:(round(v['ms_per_step'],2), round(v['achieved'] or 0)) for k,v in d['breakdown'].items()})
EOF
done
