import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib.util
from tools import gen_model, synth_audio
spec = importlib.util.spec_from_file_location("sw_binding", os.path.join(ROOT, "sentiric-stt-whisper-service_b200", "sw_binding.py"))
swb = importlib.util.module_from_spec(spec); spec.loader.exec_module(swb)
path = "/tmp/dbg_tiny_keyed.bin"
info = gen_model.generate(path, "tiny", seed=1234, script_len=40, keyed=4)
k = info["keyed"]
syms = [synth_audio.keyed_symbols(k, s) for s in range(500, 507)]
clips = [synth_audio.keyed_clip(k, sy, seed=s)[: 16000 * n] for sy, s, n in zip(syms, range(500, 507), (30, 30, 12, 30, 7, 30, 21))]
GREEDY = dict(language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=1)
def ids(r): return [t["id"] for s in r["segments"] for t in s["tokens"]]
outs = {}
for mode in ("0", "1"):
    os.environ["SW_INTERLEAVE"] = mode; os.environ["SW_INTERLEAVE_MIN"] = "1"
    e = swb.Engine(path, max_batch=4, max_beams=5, n_lanes=2)
    g = e.full_batch_pcm16(clips, e.default_params(0, **GREEDY))
    b = e.full_batch_pcm16(clips[:5], e.default_params(1, language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=1, beam_size=5))
    g2 = e.full_batch_pcm16(clips, e.default_params(0, **GREEDY))
    outs[mode] = (g, b, g2)
    e.close()
for i in range(7):
    exp = gen_model.keyed_expected_tokens(info, syms[i])
    a, b, a2, b2 = ids(outs["0"][0][i]), ids(outs["1"][0][i]), ids(outs["0"][2][i]), ids(outs["1"][2][i])
    n = min(len(a), len(b))
    diff = [j for j in range(n) if a[j] != b[j]]
    print("greedy clip", i, "len", len(a), len(b), "first diffs", diff[:5], "mode0==exp-prefix", a[:len(exp)] == exp[:len(a)], "mode1==exp-prefix", b[:len(exp)] == exp[:len(b)],
          "rerun0 same", a == a2, "rerun1 same", b == b2)
for i in range(5):
    a, b = ids(outs["0"][1][i]), ids(outs["1"][1][i])
    n = min(len(a), len(b))
    diff = [j for j in range(n) if a[j] != b[j]]
    print("beam clip", i, "len", len(a), len(b), "first diffs", diff[:5])
