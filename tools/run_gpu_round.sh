cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -x -q -s 2>&1 | tail -25 > gpurun_out/r1o_fullsize.log; cat gpurun_out/r1o_fullsize.log
