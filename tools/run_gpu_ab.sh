# A/B of the in-tree library against tools/ab_prev_librt.so (the previous commit's build) on the same box
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_vae_gpu.py tests/test_text_gpu.py -x -q 2>&1 | tail -4
PREV=$GRAFT_REPO_ROOT/tools/ab_prev_librt.so
echo "== gemm sweep prev"; RT_LIB=$PREV timeout 200 python tools/gemm_sweep.py 2,3 2>&1 | tail -9
echo "== gemm sweep new";  timeout 200 python tools/gemm_sweep.py 2,3 2>&1 | tail -9
for i in 1 2; do
  RT_LIB=$PREV timeout 400 python bench.py --no-e2e --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/ab_prev_$i.json 2> gpurun_out/ab_prev_$i.err; echo "prev $i rc=$?"
  timeout 400 python bench.py --no-e2e --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/ab_new_$i.json 2> gpurun_out/ab_new_$i.err; echo "new $i rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        b=d.get('breakdown',{})
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'gemm %.2f ms %.0f TF/s'%(b['gemm_tcgen05']['ms_per_step'],b['gemm_tcgen05']['achieved']), 'attn %.2f ms'%b['attention_tcgen05']['ms_per_step'], d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
