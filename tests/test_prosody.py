"""Prosody row: the reference's per-segment DSP (prosody_extractor.cpp:31-224) and
speaker clustering (speaker_cluster.cpp) vs (1) the CPU restatement oracle/prosody_oracle.cpp, pinned bit for
bit against golden vectors generated from the reference's OWN code (tests/golden/make_prosody_golden.py) and,
where oracle/_ref has been built, against that build live; (2) the CUDA path through the C ABI
(sw_prosody_segments_*), bit-identical to the oracle on seeded ragged segments, both input formats."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_prosody_golden import N_CASES, case  # noqa: E402
from tools import synth_audio  # noqa: E402


@pytest.fixture(scope="module")
def pros():
    from oracle import prosody
    return prosody


def test_oracle_matches_reference_golden_vectors(pros):
    g = np.load(os.path.join(ROOT, "tests", "golden", "prosody_ref.npz"))
    o = pros.oracle()
    vecs = []
    for i in range(N_CASES):
        clip, a, b, alpha = case(i)
        r = o.extract(clip[a:b], 16000, pros.default_opts(lpf_alpha=alpha))
        assert r["gender"] + ":" + r["emotion"] == str(g["tags"][i])
        got = np.array([r[f] for f in pros.FLOAT_FIELDS], np.float32)
        assert np.array_equal(got, g["floats"][i]), (i, got, g["floats"][i])  # bit-exact
        assert np.array_equal(np.array(r["speaker_vec"], np.float32), g["vecs"][i])
        vecs.append(r["speaker_vec"])
    assert o.cluster(np.array(vecs, np.float32), 0.88) == g["cluster_ids"].tolist()
    assert len(set(map(str, g["tags"]))) >= 3  # the vectors exercise several branches of the heuristics


def test_oracle_matches_reference_build_live(pros):
    ref = pros.reference()
    if ref is None:
        pytest.skip("oracle/_ref/libref_prosody.so not built here (no /root/reference)")
    o = pros.oracle()
    rng = np.random.default_rng(11)
    vecs = []
    for i in range(30):
        clip = synth_audio.to_f32(synth_audio.utterance(8, i, seconds=float(rng.uniform(0.2, 9.0))))
        seg = clip[int(rng.integers(0, len(clip) // 3)):] * np.float32(rng.choice([0.03, 0.4, 1.5]))
        opts = pros.default_opts(lpf_alpha=float(rng.choice([0.07, 0.5, 0.02])),
                                 gender_threshold=float(rng.choice([170.0, 120.0])))
        a, b = o.extract(seg, 16000, opts), ref.extract(seg, 16000, opts)
        assert a == b
        vecs.append(a["speaker_vec"])
    for edge in (np.zeros(0, np.float32), np.zeros(159, np.float32), np.zeros(160, np.float32),
                 np.full(4000, 0.3, np.float32)):
        assert o.extract(edge) == ref.extract(edge)
    assert o.extract(seg, 8000) == ref.extract(seg, 8000)
    assert o.cluster(np.array(vecs, np.float32), 0.88) == ref.cluster(np.array(vecs, np.float32), 0.88)


@pytest.mark.gpu
def test_cuda_prosody_bit_identical_to_oracle(swb, pros, micro_model):
    o = pros.oracle()
    eng = swb.Engine(micro_model[0], max_batch=2, max_beams=1, n_lanes=1)
    rng = np.random.default_rng(5)
    for u in range(6):
        secs = [30.0, 12.5, 3.0, 0.9, 30.0, 41.5][u]
        clip16 = synth_audio.utterance(10, u, seconds=secs)
        clip = synth_audio.to_f32(clip16) * np.float32([1.0, 0.3, 1.8, 0.05, 0.6, 1.0][u])
        n = len(clip)
        cuts = sorted(set([0, n] + [int(x) for x in rng.integers(0, n, size=7)]))
        segs = list(zip(cuts[:-1], cuts[1:])) + [(0, n), (n // 2, n // 2 + 100), (5, 5)]
        alpha = [0.07, 0.07, 0.2, 0.03, 0.07, 0.5][u]
        got = eng.prosody_segments(clip, segs, lpf_alpha=alpha)
        for (a, b), g in zip(segs, got):
            want = o.extract(clip[a:b], 16000, pros.default_opts(lpf_alpha=alpha))
            assert g == want, (u, a, b, {k: (g[k], want[k]) for k in g if g[k] != want[k]})
        # int16 entry: the /32768 of transcribe_pcm16 (stt_engine.cpp:117-125) happens in the load
        got16 = eng.prosody_segments(clip16, segs[:4], lpf_alpha=alpha)
        f = synth_audio.to_f32(clip16)
        for (a, b), g in zip(segs[:4], got16):
            assert g == o.extract(f[a:b], 16000, pros.default_opts(lpf_alpha=alpha))
    assert eng.prosody_segments(clip, []) == []
    with pytest.raises(RuntimeError):
        eng.prosody_segments(clip, [(10, len(clip) + 1)])
    eng.close()
