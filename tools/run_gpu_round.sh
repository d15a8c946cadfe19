cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-330 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; cut -c1-330 gpurun_out/bench_final_ref.json
