"""CPU tests of the oracle: golden vectors from the independent HF Whisper implementation
(tests/golden/make_golden.py), KV-cache consistency, logit rules, sequencing edge cases.
The reference itself holds no tests or vectors for this path."""
import ast
import os

import numpy as np
import pytest

from conftest import model_file, seg_ids
from tools import synth_audio

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "micro_hf.npz")


@pytest.fixture(scope="module")
def gold():
    g = np.load(GOLD)
    args = ast.literal_eval(str(g["model_args"]))
    path, info = model_file(args["size"], seed=args["seed"], script_len=args["script_len"])
    return g, path, info


def test_mel_matches_hf_golden(ora, gold):
    g, path, _ = gold
    o = ora.Oracle(path, act_round=ora.ACT_F32, gelu_erf=True)
    c, i = [int(v) for v in g["pcm_config"]]
    mel, n_org = o.mel(synth_audio.to_f32(synth_audio.utterance(c, i)))
    assert mel.shape == (80, 6000) and n_org == 2999
    sub = mel[::8, :3000:10]
    ref = g["hf_mel_sub"]
    # the last two frames differ by construction (upstream zero-pads, HF reflects)
    assert np.abs(sub[:, :299] - ref[:, :299]).max() < 1e-3
    # the 30 s zero pad is clamped to max-8 and normalised
    assert np.allclose(mel[:, 3100:], (mel.max() * 4 - 4 - 8 + 4) / 4, atol=1e-5)


def test_encoder_and_logits_match_hf_golden(ora, gold):
    g, path, _ = gold
    o = ora.Oracle(path, act_round=ora.ACT_F32, gelu_erf=True)
    c, i = [int(v) for v in g["pcm_config"]]
    mel, _ = o.mel(synth_audio.to_f32(synth_audio.utterance(c, i)))
    enc = o.encode(mel[:, :3000])
    ref = g["hf_enc_sub"]
    assert np.abs(enc[::15] - ref).max() / np.abs(ref).max() < 2e-4
    toks = g["tokens"]
    logits = o.decode(toks, 0)
    refl = g["hf_logits_sub"]
    assert np.abs(logits[:, g["logit_cols"]] - refl).max() / np.abs(refl).max() < 2e-4
    assert (logits.argmax(1) == g["hf_argmax"]).all()


def test_tanh_vs_erf_gelu_switch_changes_little_but_something(ora, micro_model):
    path, _ = micro_model
    o = ora.Oracle(path, act_round=ora.ACT_F32, gelu_erf=False)
    mel, _ = o.mel(synth_audio.to_f32(synth_audio.utterance(1, 1)))
    a = o.encode(mel[:, :3000])
    o2 = ora.Oracle(path, act_round=ora.ACT_F32, gelu_erf=True)
    b = o2.encode(mel[:, :3000])
    d = np.abs(a - b).max()
    assert 0 < d < 0.05


def test_incremental_decode_equals_teacher_forcing(ora, micro_model):
    path, info = micro_model
    o = ora.Oracle(path)
    mel, _ = o.mel(synth_audio.to_f32(synth_audio.utterance(1, 2)))
    o.encode(mel[:, :3000])
    sp = info["special"]
    toks = np.array([sp["sot"], sp["sot"] + 1, sp["transcribe"]] + info["script"][:9], np.int32)
    full = o.decode(toks, 0, slot=0)
    inc = np.concatenate([o.decode(toks[:3], 0, slot=1)] + [o.decode(toks[i:i + 1], i, slot=1) for i in range(3, len(toks))])
    assert np.abs(full - inc).max() < 2e-3 * np.abs(full).max()


def test_logit_rules(ora, micro_model):
    path, info = micro_model
    o = ora.Oracle(path)
    hp, sp = o.hp, info["special"]
    rng = np.random.default_rng(0)
    logits = rng.normal(0, 1, hp.n_vocab).astype(np.float32)
    p = o.default_params(0, suppress_nst=1)
    # initial position: eot, blank, specials, language tokens and timestamps > 1.00 s are suppressed
    lo, lp, pr = o.process_logits(p, [], False, 3000, 0.0, logits)
    assert lo[sp["eot"]] == -np.inf and lo[sp["sot"]] == -np.inf and lo[sp["not_"]] == -np.inf
    assert np.isinf(lo[sp["sot"] + 1: sp["sot"] + 1 + sp["n_langs"]]).all()
    assert np.isinf(lo[sp["beg"] + 51:]).all() and np.isfinite(lo[sp["beg"]: sp["beg"] + 51]).any()
    assert abs(pr.sum() - 1.0) < 1e-3 or pr[: sp["beg"]].sum() == 0  # renormalised only if text kept
    # after text + one timestamp: only timestamps / eot allowed (text suppressed)
    lo, lp, pr = o.process_logits(p, [sp["beg"], 1000, sp["beg"] + 30], True, 60, 0.0, logits)
    assert np.isinf(lo[: sp["eot"]]).all()
    assert np.isinf(lo[sp["beg"]: sp["beg"] + 30]).all()  # non-decreasing timestamps
    # after a timestamp pair: timestamps are suppressed
    lo, lp, pr = o.process_logits(p, [sp["beg"], 1000, sp["beg"] + 30, sp["beg"] + 30], True, 60, 0.0, logits)
    assert np.isinf(lo[sp["beg"]:]).all() and np.isfinite(lo[1000])
    # timestamp mass dominating any text token forces a timestamp
    big = logits.copy()
    big[sp["beg"] + 35: sp["beg"] + 45] += 12.0
    lo, lp, pr = o.process_logits(p, [sp["beg"], 1000], False, 3000, 0.0, big)
    assert pr[: sp["beg"]].sum() == 0 and pr.argmax() >= sp["beg"]
    # temperature divides the logits (a dominant text token keeps the text branch alive)
    txt = logits.copy()
    txt[2000] += 20.0
    lo1, _, _ = o.process_logits(p, [sp["beg"], 1000], False, 3000, 0.5, txt)
    assert np.isclose(lo1[2000], txt[2000] / 0.5)
    # suppress_nst off keeps the symbol tokens
    p2 = o.default_params(0, suppress_nst=0)
    a, _, _ = o.process_logits(p, [sp["beg"], 1000], False, 3000, 0.0, txt)
    b, _, _ = o.process_logits(p2, [sp["beg"], 1000], False, 3000, 0.0, txt)
    assert np.isinf(a).sum() > np.isinf(b).sum()


def test_full_follows_script_and_segments(ora, micro_model):
    path, info = micro_model
    o = ora.Oracle(path)
    p = o.default_params(0, language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=1)
    r = o.full(synth_audio.to_f32(synth_audio.utterance(1, 3)), p)
    script = info["script"]
    ids = seg_ids(r)
    # the decoder follows the scripted transcript (second timestamp of a pair and EOT are not kept)
    kept = [t for i, t in enumerate(script[:-1]) if not (i > 0 and t >= info["special"]["beg"] and script[i - 1] == t)]
    assert ids == kept
    assert r["n_windows"] == 1 and len(r["segments"]) >= 2
    segs = r["segments"]
    assert segs[0]["t0"] == 0 and segs[-1]["t1"] == 3000
    for a, b in zip(segs, segs[1:]):
        assert a["t1"] == b["t0"]
    for s in segs:
        for t in s["tokens"]:
            if t["id"] < info["special"]["eot"]:
                assert s["t0"] <= t["t0"] <= t["t1"] <= s["t1"]


def test_short_and_empty_inputs(ora, micro_model):
    path, _ = micro_model
    o = ora.Oracle(path)
    p = o.default_params(0, language="en", temperature_inc=0.0)
    assert o.full(np.zeros(0, np.float32), p)["segments"] == []
    assert o.full(np.zeros(800, np.float32), p)["n_windows"] == 0          # < 100 ms
    assert o.full(np.zeros(12000, np.float32), p)["n_windows"] == 0         # < 1 s: loop never entered
    mel, n_org = o.mel(np.zeros(0, np.float32))
    assert mel.shape == (80, 3000) and n_org == -0  # 1 + (0 + 200 - 400) // 160 in C = 0


def test_second_window_when_transcript_ends_early(ora):
    # transcript ending on a timestamp PAIR at 20.00 s: upstream seeks there and decodes a second
    # window (a single closing timestamp would skip to the end of the audio instead)
    path, info = model_file("micro", script_len=30, script_end_cs=2000, script_final_pair=True)
    o = ora.Oracle(path)
    p = o.default_params(0, language="en", temperature_inc=0.0, suppress_nst=1)
    r = o.full(synth_audio.to_f32(synth_audio.utterance(1, 4)), p)
    assert r["n_windows"] >= 2
    assert max(s["t1"] for s in r["segments"]) > 2000


def test_beam_search_is_deterministic(ora, micro_model):
    path, _ = micro_model
    o = ora.Oracle(path)
    p = o.default_params(1, language="en", temperature_inc=0.0, beam_size=3)
    pcm = synth_audio.to_f32(synth_audio.utterance(1, 5, seconds=10.0))
    assert seg_ids(o.full(pcm, p)) == seg_ids(o.full(pcm, p))


def test_tokenizer_roundtrip(ora, micro_model):
    path, _ = micro_model
    o = ora.Oracle(path)
    words = [o.token_str(i).decode() for i in (1000, 2000, 3000)]
    toks = o.tokenize("".join(words))
    assert "".join(o.token_str(int(t)).decode() for t in toks) == "".join(words)
