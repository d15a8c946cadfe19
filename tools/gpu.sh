#!/bin/bash
# One parametrised GPU-box script (replaces the per-run run_gpu_*.sh files).
#   gpurun --timeout T -- 'bash tools/gpu.sh <tag> <leg> [<leg> ...]'
# legs: tests | tests:<pytest -k expr> | bench | bench:<extra bench.py args> | ncu_list | traffic | attn_lib | py:<script and args>
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
tag=$1; shift
for leg in "$@"; do
  name=${leg%%:*}; arg=""; [[ "$leg" == *:* ]] && arg=${leg#*:}
  case $name in
    tests)
      if [ -n "$arg" ]; then
        timeout 1500 python -m pytest tests -m gpu -q -x -s -p no:warnings -k "$arg" > gpurun_out/${tag}_tests_k.txt 2>&1
        echo "tests[-k $arg] rc=$?"; tail -4 gpurun_out/${tag}_tests_k.txt | cut -c1-300
      else
        timeout 1800 python -m pytest tests -m gpu -q -x -s -p no:warnings > gpurun_out/${tag}_pytest_gpu.txt 2>&1
        echo "tests rc=$?"; tail -4 gpurun_out/${tag}_pytest_gpu.txt | cut -c1-300
      fi ;;
    bench)
      timeout 900 python bench.py $arg > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
      echo "bench rc=$?"; tail -c 3500 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err ;;
    ncu_list)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:rt:: -c 6000 --csv \
        --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-e2e \
        > gpurun_out/${tag}_ncu_list.log 2>&1
      echo "ncu_list rc=$?"; python tools/ncu_summary.py launches gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_ncu_launch_list.txt 2>&1; tail -15 gpurun_out/${tag}_ncu_launch_list.txt ;;
    traffic)
      # DRAM bytes of the tcgen05 GEMM / attention launches of ONE step (245 launches per cfg-2 step; the 15 warm-up steps
      # are skipped, i.e. run unprofiled) -> the bench line's roofline.traffic
      timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:'gemm_tc_kernel|attn_tc_kernel' --launch-skip $((15 * 245)) -c 245 --csv --log-file gpurun_out/${tag}_traffic.csv \
        python bench.py --steps 1 --warmup 15 --no-cpu-baseline --no-gpu-baseline --no-e2e > gpurun_out/${tag}_traffic.log 2>&1
      echo "traffic rc=$?"; python tools/ncu_summary.py traffic gpurun_out/${tag}_traffic.csv > gpurun_out/${tag}_gemm_dram_traffic.json 2>&1; tail -30 gpurun_out/${tag}_gemm_dram_traffic.json ;;
    attn_lib)
      timeout 900 python tools/attn_lib_compare.py > gpurun_out/${tag}_attn_lib_compare.txt 2> gpurun_out/${tag}_attn_lib_compare.err
      echo "attn_lib rc=$?"; grep -E "VERDICT|TF" gpurun_out/${tag}_attn_lib_compare.txt | tail -30 ;;
    py)
      out=gpurun_out/${tag}_$(basename ${arg%% *} .py).txt
      timeout 900 python $arg > $out 2>&1; echo "py[$arg] rc=$?"; tail -25 $out | cut -c1-300 ;;
    *) echo "unknown leg $leg" ;;
  esac
done
