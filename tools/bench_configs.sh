#!/bin/bash
# BASELINE.json configs[0..4] on one GPU box (bench.py --config 1..5), one JSON line each into
# gpurun_out/<tag>_config<N>.json; config 5 also with the facade leg (e2e_facade). Usage: tools/bench_configs.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
for c in 1 2 3 4; do
  python bench.py --config $c --steps 3 --warmup 3 --facade > gpurun_out/${tag}_config$c.json 2> gpurun_out/${tag}_config$c.err
  echo "config $c rc=$?"; tail -c 400 gpurun_out/${tag}_config$c.err
done
python bench.py --steps 3 --warmup 3 --facade > gpurun_out/${tag}_config5.json 2> gpurun_out/${tag}_config5.err
echo "config 5 rc=$?"; tail -c 400 gpurun_out/${tag}_config5.err
python - <<PY
import json
for c in range(1, 6):
    try:
        d = json.load(open("gpurun_out/${tag}_config%d.json" % c))
        print(c, d["config"]["model"], "value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), "facade", d.get("e2e_facade"),
              "parity", {k: v for k, v in d["parity_check"].items() if k != "note"}, "roofline %.3f" % d["roofline"]["frac"],
              "enc %.3f dec %.3f" % (d["stages"]["encoder_frac_of_sustained_peak"], d["stages"]["decode_frac_of_hbm"]),
              "cpu", d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(c, "failed", e)
PY
