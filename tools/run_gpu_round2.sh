cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r1o_tests.log
tail -6 gpurun_out/r1o_tests.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench16.json 2> gpurun_out/bench16.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench16.json | tr ',' '\n' | grep -E '"value"|ms_per_step|encode_ms|decode_ms|"frac"' | head -12
timeout 300 python tools/vae_once.py > gpurun_out/vae_once.log 2>&1; cat gpurun_out/vae_once.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/vae_launches.csv python tools/vae_once.py > gpurun_out/vae_ncu.log 2>&1
echo "ncu launches rc=$?"
VAE_WARMUP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 60 -c 4 -o gpurun_out/vae_gemm -f python tools/vae_once.py > gpurun_out/vae_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
