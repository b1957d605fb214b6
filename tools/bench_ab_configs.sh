#!/bin/bash
# development A/B (GPU box): BASELINE configs 1-4 short runs under environment switches
# usage: tools/bench_ab_configs.sh <tag> "VAR=val ..." "VAR=val" ...
tag=${1:-abc}; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  for c in 1 2 3 4; do
    env $envs python bench.py --config $c --steps 2 --warmup 2 --no-cpu-baseline --no-facade > gpurun_out/${tag}_${i}_c$c.json 2> gpurun_out/${tag}_${i}_c$c.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${i}_c$c.json"))
    print("[$envs] config $c value %.1f ms/step %.2f dec_frac %.3f parity %d/%d" % (d["value"], d["ms_per_step"], d["stages"]["decode_frac_of_hbm"],
          d["parity_check"]["token_identical_to_expected"], d["parity_check"]["windows"]))
except Exception as e:
    print("[$envs] config $c failed", e)
PY
  done
done
