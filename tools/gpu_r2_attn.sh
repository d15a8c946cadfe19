cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "attention" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_reference_gpu.py -q -s 2>&1 | grep -E "^ref_|passed|failed|Error" | cut -c1-300
timeout 300 python tools/attn_sweep.py 2,72,73,74,75 2>&1 | tee gpurun_out/r2_attn_pair_sweep.txt
