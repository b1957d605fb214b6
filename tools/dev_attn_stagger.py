"""Sweep of the encoder attention's CTA start offset (SW_ATTN_STAGGER, cycles) on a 2-layer model with
large-v3 widths: device ms of the encoder per batch; the attention's share is the difference / 2 layers."""
import importlib.util, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import gen_model
spec = importlib.util.spec_from_file_location("sw_binding", os.path.join(ROOT, "sentiric-stt-whisper-service_b200", "sw_binding.py"))
swb = importlib.util.module_from_spec(spec); spec.loader.exec_module(swb)
path = "/tmp/sw_lv3_2l.bin"
if not os.path.exists(path):
    gen_model.generate(path, "large-v3-2l", seed=7, script_len=20)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
values = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 500, 1000, 1500, 2000, 2500, 3000, 4000, 0]
os.environ["SW_LANES"] = "1"
eng = swb.Engine(path, max_batch=n, max_beams=1)
mel = (np.random.default_rng(0).standard_normal((n, 128, 3000)) * 0.3).astype(np.float32)
ref = None
for v in values:
    os.environ["SW_ATTN_STAGGER"] = str(v)
    for _ in range(2):
        out = eng.encode(mel, want_output=False)
    eng.stats(reset=True)
    for _ in range(5):
        eng.encode(mel, want_output=False)
    st = eng.stats(reset=True)
    ms = st["ms_encode"] / 5
    o = np.asarray(eng.encode(mel[:2], want_output=True))
    if ref is None:
        ref = o
    print("stagger %5d: encoder %.3f ms per batch of %d (2 layers)  output identical to the first: %s"
          % (v, ms, n, bool(np.array_equal(o, ref))), flush=True)
