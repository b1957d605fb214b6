#!/bin/bash
# one-off probe of the GPU box: any whisper.cpp checkout / ggml models / audio?
{
  nvidia-smi
  nproc; free -g | head -2; lscpu | head -20
  find / \( -name 'ggml-*.bin' -o -name 'whisper.h' -o -name '*.wav' -o -name 'whisper.cpp' \) -not -path '/proc/*' 2>/dev/null | head -20
  echo "probe done"
} > gpurun_out/probe.log 2>&1
