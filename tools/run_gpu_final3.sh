cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu -p no:warnings > gpurun_out/final5_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final5_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
