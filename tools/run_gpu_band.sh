cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -x -q -p no:warnings -k "gemm or forward or step" 2>&1 | tail -3
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_tc --csv --log-file gpurun_out/band_probe.csv python tools/gemm_band_probe.py > gpurun_out/band_probe.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/band_probe.csv')))
hi=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]; h=rows[hi]
idi,mn,mv=h.index("ID"),h.index("Metric Name"),h.index("Metric Value")
d={}
for r in rows[hi+1:]:
    if len(r)>mv: d.setdefault(r[idi],{})[r[mn]]=r[mv]
for k,v in d.items(): print(k, v)
PY
for i in 1 2; do
  RT_OPTIONS=gemm_band=-1 timeout 400 python bench.py --no-e2e --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/band_off_$i.json 2>/dev/null; echo "off $i rc=$?"
  timeout 400 python bench.py --no-e2e --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/band_auto_$i.json 2>/dev/null; echo "auto $i rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/band_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); b=d['breakdown']
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'gemm %.2f ms %.0f TF/s'%(b['gemm_tcgen05']['ms_per_step'],b['gemm_tcgen05']['achieved']), 'attn %.2f'%b['attention_tcgen05']['ms_per_step'], d['clocks']['sm_mhz'])
    except Exception as e: print(f,'ERR',e)
PY
