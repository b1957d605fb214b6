"""Development check (run on the GPU box): engine stages vs the CPU oracle on a small model."""
import importlib.util, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ora
from tools import gen_model, synth_audio
spec = importlib.util.spec_from_file_location("sw_binding", os.path.join(ROOT, "sentiric-stt-whisper-service_b200", "sw_binding.py"))
swb = importlib.util.module_from_spec(spec); spec.loader.exec_module(swb)

size = sys.argv[1] if len(sys.argv) > 1 else "micro"
path = "/tmp/sw_%s.bin" % size
info = gen_model.generate(path, size, seed=1234, script_len=40)
eng = swb.Engine(path, max_batch=4, max_beams=5)
o = ora.Oracle(path, weight_round=True, act_round=ora.ACT_BF16)
pcm16 = synth_audio.utterance(1, 0)
pcm = synth_audio.to_f32(pcm16)

def rel(a, b):
    return float(np.abs(a - b).max() / max(1e-9, np.abs(b).max()))

t = time.time(); g_mel = eng.mel_pcm16(pcm16); t_g = time.time() - t
o_mel, _ = o.mel(pcm)
print("mel shape", g_mel.shape, "max abs diff", float(np.abs(g_mel - o_mel).max()), "rel", rel(g_mel, o_mel))
g_mel32 = eng.mel_f32(pcm)
print("mel f32 entry max abs diff", float(np.abs(g_mel32 - o_mel).max()))
win = o_mel[:, :3000]
g_enc = eng.encode(np.stack([win, win * 0.5]))
o_enc = o.encode(win)
print("encoder rel err (bf16 oracle)", rel(g_enc[0], o_enc), "rms rel", float(np.sqrt(((g_enc[0]-o_enc)**2).mean())/o_enc.std()))
o16 = ora.Oracle(path, weight_round=False, act_round=ora.ACT_F16)
o_enc16 = o16.encode(win)
print("encoder rel err (f16 whisper.cpp-mode oracle)", rel(g_enc[0], o_enc16), "rms rel", float(np.sqrt(((g_enc[0]-o_enc16)**2).mean())/o_enc16.std()))
sp = info["special"]
toks = np.array([sp["sot"], sp["sot"] + 1, sp["transcribe"]] + info["script"][:13], np.int32)
o.encode(win)
o_log = o.decode(toks, 0)
g_log = eng.decode_logits(np.stack([toks, toks]))
print("logits rel err win0", rel(g_log[0], o_log), "argmax agree", int((g_log[0].argmax(1) == o_log.argmax(1)).sum()), "/", len(toks))
o.encode(win * 0.5)
o_log1 = o.decode(toks, 0)
print("logits rel err win1", rel(g_log[1], o_log1))
for strat, kw in ((0, {}), (1, dict(beam_size=3))):
    pe = eng.default_params(strat, language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=1, **kw)
    po = o16.default_params(strat, language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=1, **kw)
    t = time.time(); rg = eng.full_batch_pcm16([pcm16, pcm16[:16000 * 12]], pe); tg = time.time() - t
    ro = [o16.full(pcm, po), o16.full(pcm[:16000 * 12], po)]
    for a, b in zip(rg, ro):
        ia = [t["id"] for s in a["segments"] for t in s["tokens"]]
        ib = [t["id"] for s in b["segments"] for t in s["tokens"]]
        print("strategy", strat, "tokens equal", ia == ib, len(ia), len(ib), "segs", len(a["segments"]), len(b["segments"]),
              "t0/t1 equal", [(s["t0"], s["t1"]) for s in a["segments"]] == [(s["t0"], s["t1"]) for s in b["segments"]],
              "text equal", [s["text"] for s in a["segments"]] == [s["text"] for s in b["segments"]])
        if ia != ib:
            print(" gpu", ia[:60]); print(" ora", ib[:60])
        else:
            pa = np.array([t["p"] for s in a["segments"] for t in s["tokens"]]); pb = np.array([t["p"] for s in b["segments"] for t in s["tokens"]])
            ta = [(t["t0"], t["t1"]) for s in a["segments"] for t in s["tokens"]]; tb = [(t["t0"], t["t1"]) for s in b["segments"] for t in s["tokens"]]
            print("  max |p diff|", float(np.abs(pa - pb).max()) if len(pa) else 0, "token t0/t1 equal", ta == tb)
    print("  gpu time %.3fs" % tg, eng.stats())
print("DONE")
