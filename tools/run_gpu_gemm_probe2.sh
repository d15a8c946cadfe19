cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== mma shape probe"; timeout 120 python tools/mma_shape_probe.py 2>&1 | tail -6
echo "== gemm diag"; timeout 300 python tools/gemm_diag.py 2>&1 | tail -6
echo "== gelu vs bias"; timeout 120 python - <<'PY' 2>&1 | tail -8
import sys; sys.path.insert(0,'.')
from tools.gemm_sweep import bench
from reptext_b200 import _lib as L
for rep in range(2):
    for mode,name in ((L.EPI_BIAS,'bias'),(L.EPI_GELU,'gelu')):
        for dbg in (0,1):
            L.set_option("gemm_debug", dbg)
            ms,tf=bench(4608,12288,3072,3,mode=mode)
            print(f"{name} dbg{dbg}: {ms:.3f} ms {tf:.0f} TF/s", flush=True)
L.set_option("gemm_debug", 0)
PY
echo "== vae / text"; timeout 200 python tools/vae_once.py 2>&1 | tail -4 | cut -c1-400; timeout 200 python tools/text_once.py 2>&1 | tail -3 | cut -c1-400
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/gemm_probe2 -f python tools/gemm_ncu_probe2.py > gpurun_out/gemm_probe2_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/gemm_probe2_ncu.log
