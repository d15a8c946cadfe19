cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/mma_shape_probe.py > gpurun_out/mma_shape_probe.log 2>&1; cat gpurun_out/mma_shape_probe.log
