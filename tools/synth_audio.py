"""Seeded synthetic 16 kHz mono speech-like audio; there is no audio on disk.

Utterance i of config c: rng = default_rng(1000*c + i); voiced harmonic stack with vibrato,
syllabic AM, on/off bursts, white noise sigma 0.01, peak 0.5, int16.
"""
import numpy as np

SR = 16000


def utterance(config, index, seconds=30.0):
    rng = np.random.default_rng(1000 * config + index)
    n = int(round(seconds * 100)) * 160
    t = np.arange(n, dtype=np.float64) / SR
    f0 = rng.uniform(90.0, 250.0)
    vib = 1.0 + 0.03 * np.sin(2 * np.pi * 5.0 * t)
    phase = 2 * np.pi * np.cumsum(f0 * vib) / SR
    x = np.zeros(n)
    for k in range(1, 9):
        x += np.sin(k * phase) / k
    x *= np.abs(np.sin(2 * np.pi * 4.0 * t))
    gate = np.zeros(n)
    pos, on = 0, True
    while pos < n:
        ln = int(rng.uniform(0.3, 1.5) * SR)
        if on:
            gate[pos:pos + ln] = 1.0
        pos += ln
        on = not on
    x = x * gate + rng.normal(0.0, 0.01, n)
    x = 0.5 * x / np.max(np.abs(x))
    return np.round(x * 32767.0).astype(np.int16)


def durations_config4(n=256):
    """config 4: dur ~ U(5, 30) s rounded to 10 ms, seed 4000+i"""
    return [round(float(np.random.default_rng(4000 + i).uniform(5.0, 30.0)), 2) for i in range(n)]


def to_f32(pcm16):
    """the reference's transcribe_pcm16 conversion (stt_engine.cpp:117-125)"""
    return pcm16.astype(np.float32) / np.float32(32768.0)


def keyed_clip(keyed, symbols, seed=0, sr=SR, seconds=30.0, noise=0.003, amp=0.4):
    """A clip for a keyed ("listening") model (tools/gen_model.py --keyed): one tone burst per text slot, at
    the frequency of the band that names alternative symbols[i] (jittered by the seed inside the band), with
    raised-cosine edges and a little white noise. `keyed` is info["keyed"] of the generated model.
    sr != 16000 gives the same signal sampled at another rate (for the resampling front door)."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    x = rng.normal(0.0, noise, n)
    t = np.arange(n, dtype=np.float64) / sr
    for (f0, f1), sym in zip(keyed["slots"], symbols):
        a, b = int(f0 * 0.02 * sr), min(n, int(f1 * 0.02 * sr))
        if b <= a:
            continue
        f = keyed["tones_hz"][sym] * float(rng.uniform(0.97, 1.03))
        g = float(rng.uniform(0.7, 1.0)) * amp
        m = b - a
        ramp = min(int(0.02 * sr), m // 4)
        env = np.ones(m)
        if ramp > 0:
            r = 0.5 - 0.5 * np.cos(np.pi * np.arange(ramp) / ramp)
            env[:ramp] = r
            env[m - ramp:] = r[::-1]
        x[a:b] += g * env * np.sin(2 * np.pi * f * t[a:b] + float(rng.uniform(0, 2 * np.pi)))
    x = np.clip(x, -1.0, 1.0)
    return np.round(x * 32767.0).astype(np.int16) if sr == SR else x.astype(np.float32)


def keyed_symbols(keyed, seed):
    return [int(v) for v in np.random.default_rng(seed).integers(0, keyed["K"], size=len(keyed["slots"]))]
