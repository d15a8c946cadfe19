cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_ops_gpu.py -x -q -k "attention_tcgen05_half_row" 2>&1 | tail -6
if [ ${PIPESTATUS[0]} -eq 0 ]; then timeout 200 python tools/attn_sweep.py 2,62,63,64,65,2,62 2>&1 | tail -16; fi
