"""Golden vectors of the prosody row, generated from the REFERENCE's own code: oracle/_ref/libref_prosody.so
is /root/reference/src/prosody_extractor.cpp + speaker_cluster.cpp compiled where they lie (oracle/Makefile).
Run in the build container (needs /root/reference):  python tests/golden/make_prosody_golden.py
Writes tests/golden/prosody_ref.npz: for each seeded case the slice parameters and the reference's outputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import prosody  # noqa: E402
from tools import synth_audio  # noqa: E402


def case(i):
    """Seeded segment i: a slice of a synthetic utterance at a seeded gain (deterministic, no files)."""
    rng = np.random.default_rng(7000 + i)
    clip = synth_audio.to_f32(synth_audio.utterance(9, i, seconds=float(np.round(rng.uniform(0.3, 14.0), 2))))
    a = int(rng.integers(0, len(clip) // 2))
    b = int(rng.integers(a, len(clip) + 1))
    if i % 9 == 0:
        b = min(len(clip), a + int(rng.integers(0, 200)))  # around the 160-sample gate
    gain = float(rng.choice([0.02, 0.05, 0.3, 1.0, 1.9]))
    alpha = float(rng.choice([0.07, 0.07, 0.2, 0.03]))
    return clip * np.float32(gain), a, b, alpha


N_CASES = 36
if __name__ == "__main__":
    ref = prosody.reference()
    assert ref is not None, "oracle/_ref/libref_prosody.so missing: run make -C oracle (needs /root/reference)"
    floats, vecs, tags = [], [], []
    for i in range(N_CASES):
        clip, a, b, alpha = case(i)
        r = ref.extract(clip[a:b], 16000, prosody.default_opts(lpf_alpha=alpha))
        floats.append([r[f] for f in prosody.FLOAT_FIELDS])
        vecs.append(r["speaker_vec"])
        tags.append(r["gender"] + ":" + r["emotion"])
    vecs = np.array(vecs, np.float32)
    np.savez(os.path.join(ROOT, "tests", "golden", "prosody_ref.npz"), floats=np.array(floats, np.float32), vecs=vecs,
             tags=np.array(tags), cluster_ids=np.array(ref.cluster(vecs, 0.88), np.int32))
    print("wrote", N_CASES, "cases;", sorted(set(tags)))
