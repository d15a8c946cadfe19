# Full validation of the tree as the driver runs it at round end: GPU tests, smoke(), both bench arms.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=25 > gpurun_out/final_tests.log 2>&1
echo "tests rc=$?"; tail -40 gpurun_out/final_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/final_smoke.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
echo "ref rc=$?"; cut -c1-400 gpurun_out/final_bench_ref.json
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/final_bench.err | cut -c1-300; cut -c1-1500 gpurun_out/final_bench.json
timeout 600 python tools/bench_inpaint.py > gpurun_out/final_inpaint.json 2> gpurun_out/final_inpaint.err; echo "inpaint rc=$?"; tail -1 gpurun_out/final_inpaint.json | cut -c1-400
timeout 600 python bench.py --workload cfg5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/final_cfg5_1gpu.json 2> gpurun_out/final_cfg5_1gpu.err; echo "cfg5 rc=$?"; cut -c1-300 gpurun_out/final_cfg5_1gpu.json
