set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r1c_tests.log; cat gpurun_out/r1c_tests.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench7.json 2> gpurun_out/bench7.err; cut -c1-300 gpurun_out/bench7.json
RT_OPTIONS=gemm_debug=4 timeout 600 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench7_directepi.json 2>&1
timeout 300 python tools/gemm_diag.py > gpurun_out/gemm_diag2.log 2>&1; cat gpurun_out/gemm_diag2.log
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_ln.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ln_mod -s 40 -c 3 -o gpurun_out/prof_ln_cta python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ln.log 2>&1
tail -3 gpurun_out/ncu_ln.log
