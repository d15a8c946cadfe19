set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r1d_tests.log; cat gpurun_out/r1d_tests.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; cut -c1-300 gpurun_out/bench8.json
