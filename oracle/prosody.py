"""ctypes binding of the prosody oracle (oracle/prosody_oracle.h) and, when it has been built, of the
reference's own implementation (oracle/_ref/libref_prosody.so, see oracle/Makefile).

TEST INFRASTRUCTURE ONLY: imported by tests/ and bench.py's cpu_baseline leg, never by the product path."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORA_PATH = os.path.join(HERE, "build", "libprosody_oracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libref_prosody.so")
EMOTIONS = ["neutral", "excited", "sad", "angry"]


class Opts(C.Structure):
    _fields_ = [("lpf_alpha", C.c_float), ("gender_threshold", C.c_float), ("min_pitch", C.c_float),
                ("max_pitch", C.c_float)]


class Prosody(C.Structure):
    _fields_ = [("gender", C.c_char), ("emotion", C.c_int), ("arousal", C.c_float), ("valence", C.c_float),
                ("pitch_mean", C.c_float), ("pitch_std", C.c_float), ("energy_mean", C.c_float),
                ("energy_std", C.c_float), ("spectral_centroid", C.c_float), ("zero_crossing_rate", C.c_float),
                ("speaker_vec", C.c_float * 8)]


FLOAT_FIELDS = ["arousal", "valence", "pitch_mean", "pitch_std", "energy_mean", "energy_std",
                "spectral_centroid", "zero_crossing_rate"]


def default_opts(**kw):
    o = Opts(0.07, 170.0, 60.0, 500.0)  # prosody_extractor.h:20-26
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def to_dict(p):
    d = dict(gender=p.gender.decode(), emotion=EMOTIONS[p.emotion], speaker_vec=[float(x) for x in p.speaker_vec])
    for f in FLOAT_FIELDS:
        d[f] = float(getattr(p, f))
    return d


class _Lib:
    def __init__(self, path, prefix):
        self.L = C.CDLL(path)
        self.extract_fn = getattr(self.L, prefix + "_prosody_extract")
        self.extract_fn.argtypes = [C.POINTER(C.c_float), C.c_size_t, C.c_int, C.POINTER(Opts), C.POINTER(Prosody)]
        self.extract_fn.restype = None
        self.cluster_fn = getattr(self.L, prefix + "_speaker_cluster")
        self.cluster_fn.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_float, C.POINTER(C.c_int)]
        self.cluster_fn.restype = None

    def extract(self, pcm, sample_rate=16000, opts=None):
        a = np.ascontiguousarray(pcm, np.float32)
        out = Prosody()
        o = opts or default_opts()
        self.extract_fn(a.ctypes.data_as(C.POINTER(C.c_float)) if len(a) else None, len(a), sample_rate,
                        C.byref(o), C.byref(out))
        return to_dict(out)

    def cluster(self, vecs, threshold=0.88):
        v = np.ascontiguousarray(vecs, np.float32).reshape(-1, 8)
        ids = (C.c_int * len(v))()
        self.cluster_fn(v.ctypes.data_as(C.POINTER(C.c_float)), len(v), threshold, ids)
        return list(ids)


def oracle():
    if not os.path.exists(ORA_PATH) or os.path.getmtime(os.path.join(HERE, "prosody_oracle.cpp")) > os.path.getmtime(ORA_PATH):
        subprocess.check_call(["make", "-s", "-C", HERE, "build/libprosody_oracle.so"])
    return _Lib(ORA_PATH, "ora")


def reference():
    """The reference's own build, or None where it has not been compiled (no /root/reference)."""
    if not os.path.exists(REF_PATH):
        if os.path.exists("/root/reference/src/prosody_extractor.cpp"):
            subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    return _Lib(REF_PATH, "ref") if os.path.exists(REF_PATH) else None
