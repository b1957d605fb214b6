"""The keyed ("listening") synthetic model (tools/gen_model.py --keyed): its transcript is decided by the
AUDIO through mel -> conv stem -> encoder -> cross-KV -> cross attention -> logits, so token parity on it
is evidence about arithmetic, not only about sequencing (a scripted model emits the same tokens whatever it
hears). CPU part: the oracle reads the clips' symbol sequences back; the wrong audio gives other tokens.
GPU part (-m gpu): the CUDA path returns the oracle's tokens, times and probabilities for a batch of
different clips, at tiny and at large-v3 widths."""
import numpy as np
import pytest

from conftest import model_file, seg_ids
from tools import gen_model, synth_audio

GREEDY = dict(language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=1)


def clips_for(info, seeds):
    k = info["keyed"]
    syms = [synth_audio.keyed_symbols(k, s) for s in seeds]
    return syms, [synth_audio.keyed_clip(k, sy, seed=s) for sy, s in zip(syms, seeds)]


def test_oracle_reads_the_symbols_back(ora):
    path, info = model_file("tiny", script_len=40, keyed=4)
    o = ora.Oracle(path, weight_round=False, act_round=ora.ACT_F16)
    p = o.default_params(0, **GREEDY)
    syms, clips = clips_for(info, [0, 1, 2])
    got = [seg_ids(o.full(synth_audio.to_f32(c), p)) for c in clips]
    for sy, g in zip(syms, got):
        assert g == gen_model.keyed_expected_tokens(info, sy)
    assert got[0] != got[1] != got[2]                       # the audio decides
    n_text = len(info["keyed"]["slots"])
    assert n_text >= 30 and len(set(map(tuple, syms))) == 3


def test_wrong_audio_gives_other_tokens(ora):
    """Sensitivity: the same symbols rendered at 48 kHz but handed over as if they were 16 kHz (a missing
    resampler: tones at a third of their pitch, three times as slow), a clip with two bands swapped, and
    silence must NOT give the expected transcript."""
    path, info = model_file("tiny", script_len=40, keyed=4)
    k = info["keyed"]
    o = ora.Oracle(path, weight_round=False, act_round=ora.ACT_F16)
    p = o.default_params(0, **GREEDY)
    sy = synth_audio.keyed_symbols(k, 5)
    want = gen_model.keyed_expected_tokens(info, sy)
    assert seg_ids(o.full(synth_audio.to_f32(synth_audio.keyed_clip(k, sy, seed=5)), p)) == want
    slow = synth_audio.keyed_clip(k, sy, seed=5, sr=48000)[: 16000 * 30]
    assert seg_ids(o.full(slow, p)) != want
    swapped = [{0: 1, 1: 0}.get(s, s) for s in sy]
    got = seg_ids(o.full(synth_audio.to_f32(synth_audio.keyed_clip(k, swapped, seed=5)), p))
    assert got == gen_model.keyed_expected_tokens(info, swapped) and got != want
    assert seg_ids(o.full(np.zeros(16000 * 30, np.float32), p)) != want


def test_non_keyed_models_are_unchanged():
    """The generator draws its random numbers in the same order as before the keyed option existed: the
    committed HF golden vectors (micro, seed 1234) still belong to the file it writes."""
    import hashlib
    path, _ = model_file("micro")
    h = hashlib.sha256(open(path, "rb").read()).hexdigest()
    assert h == "3d46afb144dc4eee4f798c371a7d0e7fd72ba9d19b1ba0a4b978b67df18ef897"


@pytest.mark.gpu
@pytest.mark.parametrize("size,n_script", [("tiny", 40), ("large-v3-2l", 100)])
def test_keyed_greedy_identical_to_oracle(swb, ora, size, n_script):
    """Seven different clips in one batch: the CUDA path, the oracle and the clips' own symbol sequences agree
    token for token; segment / token times identical, p within 1e-2."""
    from test_gpu_parity import compare_results
    path, info = model_file(size, script_len=n_script, keyed=4)
    e = swb.Engine(path, max_batch=8, max_beams=5)
    o = ora.Oracle(path, weight_round=False, act_round=ora.ACT_F16)
    syms, clips = clips_for(info, list(range(10, 17)))
    got = e.full_batch_pcm16(clips, e.default_params(0, **GREEDY))
    po = o.default_params(0, **GREEDY)
    for i, (sy, c, g) in enumerate(zip(syms, clips, got)):
        assert seg_ids(g) == gen_model.keyed_expected_tokens(info, sy)
        if i < (7 if size == "tiny" else 2):
            compare_results(g, o.full(synth_audio.to_f32(c), po))
    assert len({tuple(seg_ids(g)) for g in got}) == 7
    e.close()


@pytest.mark.gpu
def test_keyed_wrong_audio_changes_gpu_tokens(swb):
    path, info = model_file("tiny", script_len=40, keyed=4)
    k = info["keyed"]
    e = swb.Engine(path, max_batch=4)
    sy = synth_audio.keyed_symbols(k, 21)
    want = gen_model.keyed_expected_tokens(info, sy)
    good = synth_audio.keyed_clip(k, sy, seed=21)
    slow = np.round(synth_audio.keyed_clip(k, sy, seed=21, sr=48000)[: 16000 * 30] * 32767).astype(np.int16)
    got = e.full_batch_pcm16([good, slow, np.zeros(16000 * 30, np.int16)], e.default_params(0, **GREEDY))
    assert seg_ids(got[0]) == want and seg_ids(got[1]) != want and seg_ids(got[2]) != want
    e.close()


@pytest.mark.gpu
def test_keyed_long_form_follows_the_audio_window_by_window(swb, ora):
    """A 101 s utterance = three different 30 s keyed clips + 11 s of a fourth: four windows whose `seek` advances
    by the last timestamp of the previous one. Every window must spell the symbols of ITS 30 s of audio (a window
    cut at the wrong frame, or a stale cross-KV, would spell something else), in a batch next to short utterances
    that finish early. T = 0.5 / best_of 1: above 0.5 whisper.cpp does not condition a window on the previous
    text, so the scripted positions hold in every window; the sharpened distribution puts > 0.99 on one token."""
    from test_gpu_parity import compare_results
    path, info = model_file("tiny", script_len=40, keyed=4)
    k = info["keyed"]
    seeds = (400, 401, 402)
    syms = [synth_audio.keyed_symbols(k, s) for s in seeds]
    tail = synth_audio.keyed_clip(k, syms[0], seed=403)[: 16000 * 11]
    long_clip = np.concatenate([synth_audio.keyed_clip(k, sy, seed=s) for sy, s in zip(syms, seeds)] + [tail])
    short = synth_audio.keyed_clip(k, syms[1], seed=401)[: 16000 * 9]
    kw = dict(language="en", temperature=0.5, temperature_inc=0.0, best_of=1, suppress_nst=1, token_timestamps=1)
    e = swb.Engine(path, max_batch=4)
    got = e.full_batch_pcm16([short, long_clip, short[: 16000 * 3], long_clip[: 16000 * 61]], e.default_params(0, **kw))
    want_ids = sum([gen_model.keyed_expected_tokens(info, sy) for sy in syms], [])
    ids = seg_ids(got[1])
    assert got[1]["n_windows"] == 4 and ids[: len(want_ids)] == want_ids
    ks = gen_model.keyed_sure_prefix(info, len(tail))
    assert ids[len(want_ids): len(want_ids) + ks] == gen_model.keyed_expected_tokens(info, syms[0])[:ks]
    assert [s["t0"] for s in got[1]["segments"]][::2] == [0, 3000, 6000, 9000]
    two = seg_ids(got[3])
    assert two[: 2 * len(want_ids) // 3] == want_ids[: 2 * len(want_ids) // 3]
    o = ora.Oracle(path, weight_round=False, act_round=ora.ACT_F16)
    compare_results(got[1], o.full(synth_audio.to_f32(long_clip), o.default_params(0, **kw)), p_tol=2e-2,
                    check_windows=False)
    # the 9 s neighbour: the tokens its audio decides (past its end the alternatives tie and T = 0.5 draws decide)
    ks = gen_model.keyed_sure_prefix(info, len(short))
    want_short = o.full(synth_audio.to_f32(short), o.default_params(0, **kw))
    assert ks >= 8 and seg_ids(got[0])[:ks] == seg_ids(want_short)[:ks] == gen_model.keyed_expected_tokens(info, syms[1])[:ks]
    e.close()
