cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sp_gpu.py -m gpu -x -q -k multi_process 2>&1 | tail -30 > gpurun_out/sp2_test.log; tail -5 gpurun_out/sp2_test.log; grep "rel_l2" gpurun_out/sp_worker.log | head -4; grep -v "^\*\|OMP" gpurun_out/sp_worker.log | grep -i "error\|Traceback" -A12 | head -30
