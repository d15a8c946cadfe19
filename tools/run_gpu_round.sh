cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k graph 2>&1 | tail -25
