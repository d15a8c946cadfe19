# last call of the round: full GPU suite + smoke + default bench, then the ncu launch list of one step of the same bench command
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/final4_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final4_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/final4_bench.json 2> gpurun_out/final4_bench.err; echo "bench rc=$?"; cut -c1-420 gpurun_out/final4_bench.json
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:rt:: -s 1107 -c 369 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_v3.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_launches_v3.log | cut -c1-200; wc -l gpurun_out/launches_v3.csv
