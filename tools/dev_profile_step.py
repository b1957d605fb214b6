"""ncu target: the whole hot path on a 2-layer model with large-v3 widths (same kernels and shapes
as large-v3, 1/16 of the launches): 64 x 30 s windows, greedy, short scripted transcript."""
import importlib.util, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import gen_model, synth_audio
spec = importlib.util.spec_from_file_location("sw_binding", os.path.join(ROOT, "sentiric-stt-whisper-service_b200", "sw_binding.py"))
swb = importlib.util.module_from_spec(spec); spec.loader.exec_module(swb)
path = "/tmp/sw_lv3_2l_s12.bin"
if not os.path.exists(path):
    gen_model.generate(path, "large-v3-2l", seed=7, script_len=12)
n = 64
eng = swb.Engine(path, max_batch=n, max_beams=5)
base = [synth_audio.utterance(5, i) for i in range(4)]
clips = [base[i % 4] for i in range(n)]
p = eng.default_params(0, language="en", token_timestamps=1, suppress_nst=1, temperature_inc=0.0)
for _ in range(2):
    r = eng.full_batch_pcm16(clips, p)
print("tokens", sum(len(s["tokens"]) for s in r[0]["segments"]), eng.stats())
