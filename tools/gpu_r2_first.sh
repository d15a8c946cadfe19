# round 2, first GPU call: new reference-pinned tests, Euler dt semantics of torch, attention vs libraries
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_gpu.py -q -s 2>&1 | tail -40 > gpurun_out/r2_reference_gpu.txt
tail -15 gpurun_out/r2_reference_gpu.txt
timeout 120 python tools/euler_dt_probe.py > gpurun_out/r2_euler_dt_probe.txt 2>&1; cat gpurun_out/r2_euler_dt_probe.txt
timeout 900 python tools/attn_lib_compare.py > gpurun_out/r2_attn_lib_compare.txt 2> gpurun_out/r2_attn_lib_compare.err; grep -c . gpurun_out/r2_attn_lib_compare.txt; grep VERDICT gpurun_out/r2_attn_lib_compare.txt; tail -3 gpurun_out/r2_attn_lib_compare.err
