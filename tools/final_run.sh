#!/bin/bash
# round-end verification on one GPU box: smoke, the GPU suite, the default bench line, BASELINE configs 1-4 with the facade leg
tag=${1:-final}
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_gputest.log 2>&1; tail -3 gpurun_out/${tag}_gputest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
for c in 1 2 3 4; do
  python bench.py --config $c --steps 3 --warmup 3 --facade --oracle-windows 1 > gpurun_out/${tag}_config$c.json 2> gpurun_out/${tag}_config$c.err
  echo "config $c rc=$?"
done
python - <<PY
import json
for name in ["bench"] + ["config%d" % c for c in range(1, 5)]:
    try:
        txt = [l for l in open("gpurun_out/${tag}_%s.json" % name) if l.startswith("{")][-1]
        d = json.loads(txt)
        print(name, d["config"]["model"], "value %.1f e2e %.1f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]),
              "facade", (d.get("e2e_facade") or {}).get("value"),
              "parity", {k: v for k, v in d["parity_check"].items() if k != "note"},
              "roofline %.3f alone %s" % (d["roofline"]["frac"], (d["roofline"].get("alone") or {}).get("frac")),
              "enc %.3f dec %.3f" % (d["stages"]["encoder_frac_of_sustained_peak"], d["stages"]["decode_frac_of_hbm"]),
              "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(name, "failed", e)
PY
