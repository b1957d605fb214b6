#!/bin/bash
# development A/B (GPU box): bench.py short runs under different environment switches, one summary line each
# usage: tools/bench_ab.sh <tag> "VAR=val VAR2=val" "VAR=val" ...   (use "" for the defaults)
tag=${1:-ab}; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-facade > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_$i.json"))
    print("[$envs] value %.1f ms/step %.1f decode_ms %.1f enc_ms %.1f xattn_frac %.3f (%.1f us) dec_frac %.3f parity %d/%d" % (
        d["value"], d["ms_per_step"], d["stages"]["device_ms_per_step"]["decode"], d["stages"]["device_ms_per_step"]["encode"],
        d["roofline"]["frac"], d["roofline"]["avg_launch_ms"] * 1e3, d["stages"]["decode_frac_of_hbm"], d["parity_check"]["token_identical_to_expected"], d["parity_check"]["windows"]))
except Exception as e:
    print("[$envs] failed", e)
PY
done
