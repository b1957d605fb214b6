"""Development check (GPU box): mel / encoder / teacher-forced decode bit-identical across repeats while a
second engine runs whole transcriptions from another host thread?"""
import os, sys, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import model_file
from tools import synth_audio
from tools.dev_determinism import swb

path, info = model_file("small-4l", script_len=48)
clips = [synth_audio.utterance(3, i) for i in range(16)]
tok = np.random.default_rng(1).integers(0, 50000, size=(16, 12)).astype(np.int32)
a = swb.Engine(path, max_batch=32, max_beams=5, n_lanes=1)
b = swb.Engine(path, max_batch=32, max_beams=5, n_lanes=1)
pb = b.default_params(0, language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=1)
stop = False
def hammer():
    while not stop:
        b.full_batch_pcm16(clips, pb)
mels = np.stack([a.mel_pcm16(c)[:, :3000] for c in clips])
ref_e = a.encode(mels)
ref_l = a.decode_logits(tok)
th = threading.Thread(target=hammer); th.start()
for rep in range(12):
    m2 = np.stack([a.mel_pcm16(c)[:, :3000] for c in clips])
    e2 = a.encode(mels)
    l2 = a.decode_logits(tok)
    bad = np.argwhere(l2 != ref_l)
    if len(bad):
        print("   NaNs", int(np.isnan(l2).sum()), "max abs diff", float(np.nanmax(np.abs(l2 - ref_l))),
              "windows", sorted(set(bad[:, 0].tolist())), "positions", sorted(set(bad[:, 1].tolist())), flush=True)
    print("rep", rep, "mel mismatches", int((m2 != mels).sum()), "encoder mismatches", int((e2 != ref_e).sum()),
          "windows", sorted(set(np.argwhere(e2 != ref_e)[:, 0].tolist())), "logit mismatches", int((l2 != ref_l).sum()), flush=True)
stop = True
th.join()
a.close(); b.close()
