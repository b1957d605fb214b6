"""Development check (GPU box): whole transcriptions (greedy, beam) bit-identical across repeats while a second
engine keeps the GPU busy? Prints the differing fields."""
import os, sys, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import model_file
from tools import synth_audio
from tools.dev_determinism import swb, diff

path, info = model_file("tiny")
a = swb.Engine(path, max_batch=16, max_beams=5, n_lanes=1)
b = swb.Engine(path, max_batch=16, max_beams=5, n_lanes=1)
clips = [synth_audio.utterance(7, i, seconds=10.0 + i) for i in range(12)]
pb = b.default_params(0, language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=1)
stop = []
def hammer():
    while not stop:
        b.full_batch_pcm16(clips, pb)
cfgs = [(1, dict(language="en", temperature_inc=0.0, suppress_nst=1, beam_size=5)),
        (0, dict(language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=1)),
        (0, dict(language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=0))]
refs = [a.full_batch_pcm16(clips[:6], a.default_params(s, **kw)) for s, kw in cfgs]
th = threading.Thread(target=hammer); th.start()
for rep in range(6):
    for (s, kw), ref in zip(cfgs, refs):
        got = a.full_batch_pcm16(clips[:6], a.default_params(s, **kw))
        d = diff(ref, got)
        print("strategy", s, "tts", kw.get("token_timestamps"), "rep", rep, "diffs", len(d), [x[:90] for x in d[:3]], flush=True)
stop.append(1); th.join()
a.close(); b.close()
