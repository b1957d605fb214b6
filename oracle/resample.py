"""ctypes binding of the sample-rate-conversion oracle (oracle/resample_oracle.cpp). TEST INFRASTRUCTURE ONLY.
PARITY UNPINNED against libsamplerate (see the header of resample_oracle.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "build", "libresample_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or os.path.getmtime(os.path.join(HERE, "resample_oracle.cpp")) > os.path.getmtime(LIB_PATH):
            subprocess.check_call(["make", "-s", "-C", HERE, "build/libresample_oracle.so"])
        _lib = C.CDLL(LIB_PATH)
        _lib.ora_resample_out_len.restype = C.c_int64
        _lib.ora_resample_out_len.argtypes = [C.c_int64, C.c_int, C.c_int]
        _lib.ora_resample_f32.argtypes = [C.POINTER(C.c_float), C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_float)]
    return _lib


def resample(x, sr_in, sr_out=16000):
    L = lib()
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty(L.ora_resample_out_len(len(x), sr_in, sr_out), np.float32)
    if len(x):
        L.ora_resample_f32(x.ctypes.data_as(C.POINTER(C.c_float)), len(x), sr_in, sr_out,
                           y.ctypes.data_as(C.POINTER(C.c_float)))
    return y
