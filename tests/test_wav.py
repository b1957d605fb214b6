"""WAV front door: host/wav.h restates the reference's parse_wav_robust
(utils.h:101-202). Pinned against the reference's own code where it has been compiled
(oracle/_ref/libref_wav.so, see oracle/Makefile) on generated containers, and against properties that do not
need it (mono pass-through, stereo mix, channel pick, chunk walk, raw-PCM fallback, rejections)."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import PKG

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bind(path, name):
    L = C.CDLL(path)
    fn = getattr(L, name)
    fn.restype = C.c_long
    fn.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int16), C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]

    def parse(data):
        cap = len(data) // 2 + 16
        out = (C.c_int16 * cap)()
        sr, ch = C.c_int(0), C.c_int(0)
        n = fn(data, len(data), out, cap, C.byref(sr), C.byref(ch))
        if n < 0:
            return None
        return np.array(out[:n], np.int16), sr.value, ch.value
    return parse


@pytest.fixture(scope="module")
def ours():
    subprocess.check_call(["make", "-s", "-C", PKG])
    subprocess.check_call(["make", "-s", "-C", os.path.join(PKG, "host")])
    return _bind(os.path.join(PKG, "libstt_engine.so"), "stt_parse_wav")


@pytest.fixture(scope="module")
def ref():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_wav.so")
    if not os.path.exists(path) and os.path.exists("/root/reference/src/utils.h"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    return _bind(path, "ref_parse_wav") if os.path.exists(path) else None


def wav(samples, channels=1, rate=16000, bits=16, tag=1, extra=b"", fmt_size=16, pad_odd=False, data_size=None):
    payload = np.asarray(samples, np.int16).tobytes()
    fmt = struct.pack("<HHIIHH", tag, channels, rate, rate * channels * bits // 8, channels * bits // 8, bits)
    fmt += b"\0" * (fmt_size - 16)
    body = b"WAVE" + b"fmt " + struct.pack("<I", fmt_size) + fmt + extra
    body += b"data" + struct.pack("<I", len(payload) if data_size is None else data_size) + payload
    return b"RIFF" + struct.pack("<I", len(body)) + body


def cases():
    rng = np.random.default_rng(3)
    mono = rng.integers(-30000, 30000, 1000).astype(np.int16)
    st = rng.integers(-32768, 32767, 2001).astype(np.int16)  # odd count: the last sample has no partner
    yield "mono", wav(mono)
    yield "stereo", wav(st, channels=2, rate=44100)
    yield "three channels", wav(rng.integers(-100, 100, 999).astype(np.int16), channels=3, rate=48000)
    yield "extensible tag", wav(mono, tag=0xFFFE, fmt_size=40)
    yield "LIST chunk before data", wav(mono, extra=b"LIST" + struct.pack("<I", 10) + b"INFOabcdef")
    yield "odd chunk is padded", wav(mono, extra=b"junk" + struct.pack("<I", 3) + b"abc" + b"\0")
    yield "data size larger than the file", wav(mono, data_size=10 ** 6)
    yield "no header: raw pcm", mono.tobytes() + b"\x01"
    yield "empty", b""
    yield "8 bit", wav(mono, bits=8)
    yield "float tag", wav(mono, tag=3)
    yield "no data chunk", wav(mono)[:36]
    yield "short fmt", b"RIFF" + struct.pack("<I", 30) + b"WAVEfmt " + struct.pack("<I", 8) + b"\0" * 8 + b"data" + struct.pack("<I", 4) + b"abcd"


def test_restatement_properties(ours):
    c = dict(cases())
    pcm, sr, ch = ours(c["mono"])
    assert (sr, ch, len(pcm)) == (16000, 1, 1000)
    pcm, sr, ch = ours(c["stereo"])
    assert (sr, ch, len(pcm)) == (44100, 2, 1000)
    raw = np.frombuffer(c["stereo"][44:44 + 4000], np.int16).astype(np.int32)
    assert np.array_equal(pcm, ((raw[0::2] + raw[1::2]) / 2).astype(np.int32).astype(np.int16))  # C division truncates
    pcm, sr, ch = ours(c["three channels"])
    assert (sr, ch, len(pcm)) == (48000, 3, 333)
    assert ours(c["no header: raw pcm"])[1:] == (16000, 1) and len(ours(c["no header: raw pcm"])[0]) == 1000
    assert len(ours(c["empty"])[0]) == 0
    # a data chunk that claims more bytes than the file holds ends the chunk walk before it is accepted
    # (utils.h:153 precedes :165), exactly as in the reference
    for bad in ("8 bit", "float tag", "no data chunk", "short fmt", "data size larger than the file"):
        assert ours(c[bad]) is None, bad
    for ok in ("extensible tag", "LIST chunk before data", "odd chunk is padded"):
        assert len(ours(c[ok])[0]) == 1000, ok


def test_malformed_headers_the_reference_crashes_on_are_rejected(ours):
    """A 0-channel header divides by zero in the reference (utils.h:194-199: SIGFPE, so it cannot be a parity
    case); a 0 Hz / negative rate only fails downstream. The restatement rejects both like any other bad file."""
    mono = np.arange(100, dtype=np.int16)
    assert ours(wav(mono, channels=0)) is None
    assert ours(wav(mono, rate=0)) is None
    hdr = bytearray(wav(mono))
    hdr[24:28] = struct.pack("<i", -8000)
    assert ours(bytes(hdr)) is None
    assert len(ours(wav(mono, channels=1, rate=8000))[0]) == 100


def test_restatement_matches_the_reference_build(ours, ref):
    if ref is None:
        pytest.skip("oracle/_ref/libref_wav.so not built here (no /root/reference)")
    for name, data in cases():
        if name.startswith("no header") or name == "empty":
            continue  # the reference shells out to ffmpeg first (not in this image): covered by the properties
        a, b = ours(data), ref(data)
        assert (a is None) == (b is None), name
        if a is not None:
            assert a[1:] == b[1:] and np.array_equal(a[0], b[0]), name
