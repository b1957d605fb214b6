"""Development check (GPU box): does a partially filled batch depend on what the buffers held before?"""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import model_file
from tools import synth_audio
from tools.dev_determinism import diff, swb  # noqa

path, info = model_file("small-4l", script_len=48)
clips = [synth_audio.utterance(3, i) for i in range(32)]
kw = dict(language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=1)
for nclip in (16, 32, 5):
    e = swb.Engine(path, max_batch=32, max_beams=5, n_lanes=1)
    pe = e.default_params(0, **kw)
    ref = e.full_batch_pcm16(clips[:nclip], pe)
    for rep in range(3):
        got = e.full_batch_pcm16(clips[:nclip], pe)
        d = diff(ref, got)
        print("BN", os.environ.get("SW_SKINNY_BN"), "lanes 1 clips", nclip, "rep", rep, "diffs", len(d), d[:2], flush=True)
    e.close()
