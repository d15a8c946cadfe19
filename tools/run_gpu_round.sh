cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
RT_OPTIONS=attn_variant=52 timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k attention 2>&1 | tail -5 > gpurun_out/r1m_tests.log; cat gpurun_out/r1m_tests.log
timeout 200 python tools/attn_sweep.py 2,42,49,54,55,56,57 > gpurun_out/attn_sweep10.log 2>&1; cat gpurun_out/attn_sweep10.log
