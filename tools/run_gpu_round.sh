cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
K='regex:gemm_tc|attn_tc|ln_mod|gemv|euler|rope_table|time_sinusoid|silu_f32|cast_to_f32|gemm_simt|attn_simt'
$CMD > gpurun_out/plain_r1i.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 1107 --launch-count 369 --csv --log-file gpurun_out/launches_v2.csv $CMD > gpurun_out/ncu_launches_v2.log 2>&1
tail -1 gpurun_out/ncu_launches_v2.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 700 --launch-count 3 -o gpurun_out/prof_gemm_v2 -f $CMD > gpurun_out/ncu_gemm_v2.log 2>&1
tail -1 gpurun_out/ncu_gemm_v2.log
