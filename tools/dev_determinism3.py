"""Development check (GPU box): teacher-forced decoder logits bit-identical across repeats while a second
engine keeps the GPU busy from another host thread (what a second lane does)?"""
import os, sys, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import model_file
from tools.dev_determinism import swb

path, info = model_file("small-4l", script_len=48)
tok = np.random.default_rng(1).integers(0, 50000, size=(16, 24)).astype(np.int32)
a = swb.Engine(path, max_batch=32, max_beams=5, n_lanes=1)
b = swb.Engine(path, max_batch=32, max_beams=5, n_lanes=1)
stop = False
def hammer():
    while not stop:
        b.decode_logits(tok)
ref = a.decode_logits(tok)
for busy in (False, True):
    if busy:
        th = threading.Thread(target=hammer); th.start()
    for rep in range(6):
        got = a.decode_logits(tok)
        bad = np.argwhere(got != ref)
        print("busy", busy, "rep", rep, "mismatching logits", len(bad), "first", bad[:1].tolist(),
              "positions", sorted(set(bad[:, 1].tolist()))[:10], flush=True)
stop = True
th.join()
a.close(); b.close()
