cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_text_gpu.py tests/test_vae_gpu.py -q -s 2>&1 | tail -40 > gpurun_out/r1p_tests.log
grep -E "rel-L2|passed|failed|Error|assert" gpurun_out/r1p_tests.log | cut -c1-220 | tail -30
timeout 300 python tools/vae_once.py > gpurun_out/vae_once2.log 2>&1; cut -c1-600 gpurun_out/vae_once2.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench17.json 2> gpurun_out/bench17.err
echo "bench rc=$?"; tail -5 gpurun_out/bench17.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench17.json') if l.startswith('{')][-1])
    for k in ('value','ms_per_step','e2e','vae','text_encoders','gpu_launches'): print(k, d[k])
except Exception as e: print("no bench line", e)
PY
