set -x
for e in 0 1; do SW_ATTN_EARLY=$e timeout 200 python tools/dev_attn_stagger.py 64 0,1500,0 > gpurun_out/attn_elect_early$e.log 2>&1; done
SW_ATTN_TRACE=1 SW_ATTN_STAGGER=0 SW_ATTN_EARLY=0 SW_LANES=1 timeout 100 python tools/dev_encode_ncu.py 64 > gpurun_out/attn_trace_elect_early0.log 2>&1
SW_ATTN_TRACE=1 SW_ATTN_STAGGER=0 SW_ATTN_EARLY=1 SW_LANES=1 timeout 100 python tools/dev_encode_ncu.py 64 > gpurun_out/attn_trace_elect_early1.log 2>&1
timeout 200 python tools/dev_gemm_check.py > gpurun_out/gemm_check_elect.log 2>&1
SW_LANES=1 SW_PDL=0 timeout 200 python tools/dev_step_time.py > gpurun_out/step_time_elect.log 2>&1
SW_ATTN_EARLY=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "encoder or logits or greedy or determin" > gpurun_out/t_elect_early0.log 2>&1
SW_ATTN_EARLY=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "encoder or logits or greedy" > gpurun_out/t_elect_early1.log 2>&1
tail -4 gpurun_out/attn_elect_early*.log gpurun_out/step_time_elect.log gpurun_out/t_elect_early*.log; tail -12 gpurun_out/gemm_check_elect.log
