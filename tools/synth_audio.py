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
