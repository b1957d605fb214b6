"""Sample-rate conversion in front of the path (reference: libsamplerate through
stt_engine.cpp:87-115, NOT in the tree - parity with it is not claimed). The CPU statement of the engine's own
converter is checked for quality (tone SNR, agreement with scipy's polyphase resampler, output length), and the
CUDA kernel behind sw_resample_f32 is bit-identical to that statement."""
import numpy as np
import pytest

from tools import synth_audio


def tone(f, sr, secs=1.5, amp=0.5):
    return (amp * np.sin(2 * np.pi * f * np.arange(int(sr * secs)) / sr)).astype(np.float32)


@pytest.mark.parametrize("sr_in", [8000, 11025, 22050, 44100, 48000])
def test_oracle_tone_snr_and_length(sr_in):
    from oracle import resample
    for f in (220.0, 1000.0, 3100.0, 6000.0):
        if f > 0.4 * min(sr_in, 16000):
            continue
        x = tone(f, sr_in)
        y = resample.resample(x, sr_in)
        assert len(y) == (len(x) * 16000) // sr_in
        ref = 0.5 * np.sin(2 * np.pi * f * np.arange(len(y)) / 16000)
        m = slice(400, len(y) - 400)
        snr = 10 * np.log10(np.mean(ref[m] ** 2) / np.mean((y[m] - ref[m]) ** 2))
        assert snr > 80.0, (sr_in, f, snr)


def test_oracle_rejects_what_aliases_and_agrees_with_scipy():
    from scipy.signal import resample_poly
    from oracle import resample
    # a 12 kHz tone at 48 kHz is above the 8 kHz Nyquist of the output: it must be gone (>= 70 dB down)
    y = resample.resample(tone(12000.0, 48000), 48000)
    assert 20 * np.log10(np.abs(y[400:-400]).max() / 0.5) < -70.0
    # speech-like audio: same result as scipy's polyphase resampler up to the difference of the two filters
    clip = synth_audio.to_f32(synth_audio.utterance(11, 0, seconds=3.0))
    up = resample_poly(clip.astype(np.float64), 3, 1)  # a 48 kHz version of the clip
    y = resample.resample(up.astype(np.float32), 48000)
    n = min(len(y), len(clip))
    err = y[200:n - 200] - clip[200:n - 200]
    assert np.sqrt(np.mean(err ** 2)) < 2e-3 * np.sqrt(np.mean(clip ** 2)) + 1e-4
    assert len(resample.resample(np.zeros(0, np.float32), 44100)) == 0


@pytest.mark.gpu
def test_cuda_resampler_bit_identical_to_oracle(swb, micro_model):
    from oracle import resample
    eng = swb.Engine(micro_model[0], max_batch=2, max_beams=1, n_lanes=1)
    rng = np.random.default_rng(2)
    for sr_in in (8000, 11025, 22050, 44100, 48000, 16000, 96000):
        x = (rng.standard_normal(int(sr_in * 0.7) + 13) * 0.2).astype(np.float32) + tone(440.0, sr_in, 0.7)[0] * 0
        got = eng.resample(x, sr_in)
        want = resample.resample(x, sr_in)
        assert got.shape == want.shape and np.array_equal(got, want), (sr_in, np.abs(got - want).max())
    # up-conversion (8 kHz telephone audio) and a whole transcription of 48 kHz audio through the resampler
    assert np.array_equal(eng.resample(x[:5000], 8000, 48000), resample.resample(x[:5000], 8000, 48000))
    clip = synth_audio.to_f32(synth_audio.utterance(3, 4, seconds=8.0))
    from scipy.signal import resample_poly
    clip48 = resample_poly(clip.astype(np.float64), 3, 1).astype(np.float32)
    back = eng.resample(clip48, 48000)
    pe = eng.default_params(0, language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=1)
    a = eng.full_f32(clip, pe)
    b = eng.full_f32(back[:len(clip)], pe)
    assert [t["id"] for s in a["segments"] for t in s["tokens"]] == [t["id"] for s in b["segments"] for t in s["tokens"]]
    eng.close()
