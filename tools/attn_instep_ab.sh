#!/bin/bash
# In-step A/B of the attention kernel's exponential split (A/B build): bench.py's device-timed loop per variant.
#   gpurun -- 'bash tools/attn_instep_ab.sh "0 7 12 10"'
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for rep in 1 2; do
for v in $1; do
RT_LIB=$PWD/reptext_b200/csrc/librt_reptext_ab.so RT_OPTIONS=attn_variant=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e > gpurun_out/attn_ab.json 2>gpurun_out/attn_ab.err
python - <<P
import json
d=json.loads(open("gpurun_out/attn_ab.json").read().strip().splitlines()[-1])
print("variant $v", round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms_per_step"],2) for k,v in d["breakdown"].items()})
P
done; done
