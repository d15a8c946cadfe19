cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -p no:warnings > gpurun_out/final7_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/final7_tests.log
timeout 100 python tools/text_once.py 2>&1 | tail -3 | cut -c1-330
