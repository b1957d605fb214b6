"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the LAST occurrence of the
sampling decoder step (from embed_tokens_kernel to process_logits_kernel) grouped by kernel and grid,
and the encoder pass before it. Usage: python tools/summarize_launches.py launches.csv"""
import csv, re, sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    name = re.sub(r"^.*?::", "", r["Kernel Name"]).split("(")[0]
    rows.append((name, r["Grid Size"], float(r["Metric Value"]) / 1e3))

def table(seg, title):
    agg = OrderedDict()
    for n, g, us in seg:
        k = (n, g)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(v[1] for v in agg.values())
    print("%s: launches=%d total=%.3f ms" % (title, len(seg), tot / 1e3))
    for (n, g), (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-40s grid=%-16s n=%4d %9.3f ms %5.1f%% avg %8.1f us" % (n, g, c, us / 1e3, 100 * us / tot, us / c))

ends = [i for i, r in enumerate(rows) if r[0].startswith("process_logits_kernel")]
starts = [i for i, r in enumerate(rows) if r[0].startswith("embed_tokens_kernel")]
if ends and starts:
    e = ends[-1]
    s = max(i for i in starts if i < e)
    table(rows[s:e + 1], "one sampling decoder step (last of the run)")
    if len(sys.argv) > 2:
        for n, g, us in rows[s:e + 1]:
            print("    %-40s %-16s %8.1f" % (n, g, us))
mel = [i for i, r in enumerate(rows) if r[0].startswith("mel_log_power")]
if mel and starts:
    s = mel[-1]
    e = min(i for i in starts if i > s)
    table(rows[s:e], "front end + encoder + cross-KV (last batch)")
