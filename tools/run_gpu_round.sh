cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
RT_OPTIONS=sp_replicate_mod=1 timeout 900 python -m pytest tests/test_sp_gpu.py -m gpu -x -q 2>&1 | tail -3
