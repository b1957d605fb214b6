#!/bin/bash
# Development: bench.py under several engine switches in one GPU-box call. Usage: tools/bench_matrix.sh "ENV1=a ENV2=b" "ENV1=c" ...
# Writes one line per configuration to gpurun_out/bench_matrix.log
out=gpurun_out/bench_matrix.log
: > $out
for cfg in "$@"; do
  echo "== $cfg" >> $out
  env $cfg python bench.py --no-cpu-baseline --steps 2 --warmup 1 2>gpurun_out/bench_matrix.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['stages']
print(json.dumps(dict(value=round(d['value'],1), e2e=round(d['e2e']['value'],1), ms_per_step=round(d['ms_per_step'],1), xattn_frac=round(d['roofline']['frac'],3), xattn_ms=round(d['roofline']['avg_launch_ms'],4), dev_ms=s['device_ms_per_step'], decode_frac=round(s['decode_frac_of_hbm'],3), enc_frac=round(s['encoder_frac_of_sustained_peak'],3))))
" >> $out 2>&1
done
cat $out
