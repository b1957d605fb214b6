"""ncu target: encoder over a few windows of a 2-layer model with large-v3 widths."""
import importlib.util, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import gen_model
spec = importlib.util.spec_from_file_location("sw_binding", os.path.join(ROOT, "sentiric-stt-whisper-service_b200", "sw_binding.py"))
swb = importlib.util.module_from_spec(spec); spec.loader.exec_module(swb)
path = "/tmp/sw_lv3_2l.bin"
if not os.path.exists(path):
    gen_model.generate(path, "large-v3-2l", seed=7, script_len=20)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
eng = swb.Engine(path, max_batch=n, max_beams=1)
mel = (np.random.default_rng(0).standard_normal((n, 128, 3000)) * 0.3).astype(np.float32)
for _ in range(3):
    eng.encode(mel, want_output=False)
eng.stats(reset=True)
t = time.time()
for _ in range(5):
    eng.encode(mel, want_output=False)
st = eng.stats()
print("encode device ms per window per layer-pair model:", st["ms_encode"] / st["n_windows"], "windows", st["n_windows"])
