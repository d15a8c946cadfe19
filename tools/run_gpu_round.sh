set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1b_tests.log; cat gpurun_out/r1b_tests.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; cat gpurun_out/bench6.json | cut -c1-600
RT_OPTIONS=ln_warp_rows=1,gemv_single_row=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench6_oldln.json 2>&1
timeout 300 python tools/gemm_diag.py > gpurun_out/gemm_diag.log 2>&1; cat gpurun_out/gemm_diag.log
timeout 300 python tools/attn_sweep.py 2,6,7,8,9,10 > gpurun_out/attn_sweep.log 2>&1; cat gpurun_out/attn_sweep.log
