cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 420 python tools/vae_diag.py --size 1024 --reps 5 > gpurun_out/vae_diag.log 2>&1
echo "diag rc=$?"
grep -E "^\[(ok|FAIL)" gpurun_out/vae_diag.log | cut -c1-200 | tail -40
timeout 300 python -m pytest tests/test_vae_gpu.py -q 2>&1 | tail -25 > gpurun_out/vae_tests.log
tail -15 gpurun_out/vae_tests.log
