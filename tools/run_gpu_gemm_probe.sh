cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/gemm_ncu_probe.py > gpurun_out/gemm_probe_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/gemm_probe_plain.log
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/gemm_probe -f python tools/gemm_ncu_probe.py > gpurun_out/gemm_probe_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/gemm_probe_ncu.log
ls -la gpurun_out/*.ncu-rep
