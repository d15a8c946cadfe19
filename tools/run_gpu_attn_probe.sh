cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/attn_probe -f python tools/attn_ncu_probe.py > gpurun_out/attn_probe_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/attn_probe_ncu.log
