cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_text_gpu.py -q -s 2>&1 | tail -40 > gpurun_out/r1q_tests.log
grep -E "rel-L2|passed|failed|Error|assert" gpurun_out/r1q_tests.log | cut -c1-220 | tail -30
timeout 300 python tools/text_once.py > gpurun_out/text_once.log 2>&1; cut -c1-700 gpurun_out/text_once.log | tail -6
