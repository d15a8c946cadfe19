# round 2, second GPU call: whole GPU suite (new tests included), bench with the new legs, attention vs libraries
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2_pytest_gpu.txt 2>&1; tail -5 gpurun_out/r2_pytest_gpu.txt
grep -E "^ref_|preparation" gpurun_out/r2_pytest_gpu.txt | cut -c1-400
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -c 3000 gpurun_out/r2_bench_1gpu.json; tail -5 gpurun_out/r2_bench_1gpu.err
timeout 600 python tools/attn_lib_compare.py > gpurun_out/r2_attn_lib_compare.txt 2> gpurun_out/r2_attn_lib_compare.err; grep VERDICT gpurun_out/r2_attn_lib_compare.txt
