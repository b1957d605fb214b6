"""Development check (GPU box): does a teacher-forced sw_decode_logits call change a later beam-search run?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import model_file
from tools import synth_audio
from tools.dev_determinism import swb, diff
path, info = model_file("tiny")
a = swb.Engine(path, max_batch=16, max_beams=5, n_lanes=1)
clips = [synth_audio.utterance(7, i, seconds=10.0 + i) for i in range(12)]
tok = np.random.default_rng(3).integers(0, 50000, size=(16, 16)).astype(np.int32)
kw = dict(language="en", temperature_inc=0.0, suppress_nst=1, beam_size=5)
ref = a.full_batch_pcm16(clips[:6], a.default_params(1, **kw))
for rep in range(4):
    if rep % 2 == 1:
        a.decode_logits(tok)
    got = a.full_batch_pcm16(clips[:6], a.default_params(1, **kw))
    d = diff(ref, got)
    print("rep", rep, "after decode_logits" if rep % 2 else "plain", "diffs", len(d), [x[:100] for x in d[:4]], flush=True)
a.close()
