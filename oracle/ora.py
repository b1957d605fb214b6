"""ctypes binding of the CPU oracle (oracle/whisper_oracle.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product path never imports this module.
PARITY UNPINNED - see the header of whisper_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "build", "libwhisper_oracle.so")

ACT_F32, ACT_F16, ACT_BF16 = 0, 1, 2


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(LIB_PATH)
            for f in ("whisper_oracle.cpp", "whisper_oracle.h")):
        subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB_PATH


class HParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "n_vocab", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer", "n_text_ctx",
        "n_text_state", "n_text_head", "n_text_layer", "n_mels", "ftype", "token_eot", "token_sot",
        "token_translate", "token_transcribe", "token_solm", "token_prev", "token_nosp", "token_not",
        "token_beg", "is_multilingual")]


class FullParams(C.Structure):
    _fields_ = [
        ("strategy", C.c_int), ("beam_size", C.c_int), ("best_of", C.c_int),
        ("temperature", C.c_float), ("temperature_inc", C.c_float),
        ("entropy_thold", C.c_float), ("logprob_thold", C.c_float), ("no_speech_thold", C.c_float),
        ("translate", C.c_int), ("tdrz_enable", C.c_int), ("suppress_nst", C.c_int),
        ("suppress_blank", C.c_int), ("token_timestamps", C.c_int), ("no_timestamps", C.c_int),
        ("single_segment", C.c_int), ("no_context", C.c_int),
        ("max_initial_ts", C.c_float), ("length_penalty", C.c_float),
        ("language", C.c_char_p), ("initial_prompt", C.c_char_p),
        ("prompt_tokens", C.POINTER(C.c_int32)), ("prompt_n_tokens", C.c_int),
        ("max_tokens", C.c_int)]


class TokenData(C.Structure):
    _fields_ = [("id", C.c_int32), ("tid", C.c_int32), ("p", C.c_float), ("plog", C.c_float),
                ("pt", C.c_float), ("ptsum", C.c_float), ("t0", C.c_int64), ("t1", C.c_int64),
                ("t_dtw", C.c_int64), ("vlen", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    vp, ci, fp = C.c_void_p, C.c_int, C.POINTER(C.c_float)
    L.ora_load.restype = vp
    L.ora_load.argtypes = [C.c_char_p, ci]
    L.ora_free.argtypes = [vp]
    L.ora_last_error.restype = C.c_char_p
    L.ora_set_act_round.argtypes = [vp, ci]
    L.ora_set_gelu_erf.argtypes = [vp, ci]
    L.ora_set_threads.argtypes = [vp, ci]
    L.ora_get_hparams.argtypes = [vp, C.POINTER(HParams)]
    L.ora_token_to_str.restype = C.c_char_p
    L.ora_token_to_str.argtypes = [vp, ci]
    L.ora_lang_id.argtypes = [C.c_char_p]
    L.ora_tokenize.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int32), ci]
    L.ora_mel.argtypes = [vp, fp, ci, fp, C.POINTER(ci), C.POINTER(ci)]
    L.ora_encode.argtypes = [vp, fp, fp]
    L.ora_encode_tap.argtypes = [vp, ci, fp]
    L.ora_decode.argtypes = [vp, ci, C.POINTER(C.c_int32), ci, ci, fp]
    L.ora_full_default_params.restype = FullParams
    L.ora_full_default_params.argtypes = [ci]
    L.ora_full.argtypes = [vp, C.POINTER(FullParams), fp, ci, C.POINTER(vp)]
    L.ora_result_n_segments.argtypes = [vp]
    L.ora_result_segment_text.restype = C.c_char_p
    L.ora_result_segment_text.argtypes = [vp, ci]
    L.ora_result_segment_t0.restype = C.c_int64
    L.ora_result_segment_t0.argtypes = [vp, ci]
    L.ora_result_segment_t1.restype = C.c_int64
    L.ora_result_segment_t1.argtypes = [vp, ci]
    L.ora_result_segment_speaker_turn_next.argtypes = [vp, ci]
    L.ora_result_n_tokens.argtypes = [vp, ci]
    L.ora_result_token_data.restype = TokenData
    L.ora_result_token_data.argtypes = [vp, ci, ci]
    for n in ("lang_id", "n_decode_steps", "n_windows"):
        getattr(L, "ora_result_" + n).argtypes = [vp]
    for n in ("ms_mel", "ms_encode", "ms_decode"):
        getattr(L, "ora_result_" + n).argtypes = [vp]
        getattr(L, "ora_result_" + n).restype = C.c_double
    L.ora_result_free.argtypes = [vp]
    L.ora_process_logits.argtypes = [vp, C.POINTER(FullParams), C.POINTER(C.c_int32), ci, ci, ci,
                                     C.c_float, fp, fp, fp, fp]
    _lib = L
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def result_to_dict(L, r, prefix, tok_struct_to_dict):
    segs = []
    g = lambda n: getattr(L, prefix + n)
    for i in range(g("result_n_segments")(r)):
        toks = [tok_struct_to_dict(g("result_token_data")(r, i, j))
                for j in range(g("result_n_tokens")(r, i))]
        segs.append(dict(text=g("result_segment_text")(r, i), t0=g("result_segment_t0")(r, i),
                         t1=g("result_segment_t1")(r, i),
                         speaker_turn_next=bool(g("result_segment_speaker_turn_next")(r, i)),
                         tokens=toks))
    return segs


def tok_to_dict(t):
    return dict(id=t.id, tid=t.tid, p=t.p, plog=t.plog, pt=t.pt, ptsum=t.ptsum, t0=t.t0, t1=t.t1,
                vlen=t.vlen)


class Oracle:
    """One loaded model. weight_round=True rounds matrix weights to bf16 (what the engine holds)."""

    def __init__(self, path, weight_round=False, act_round=ACT_F16, gelu_erf=False, threads=0):
        self.L = lib()
        self.h = self.L.ora_load(path.encode(), int(weight_round))
        if not self.h:
            raise RuntimeError("ora_load: " + self.L.ora_last_error().decode())
        self.L.ora_set_act_round(self.h, act_round)
        self.L.ora_set_gelu_erf(self.h, int(gelu_erf))
        self.L.ora_set_threads(self.h, threads)
        hp = HParams()
        self.L.ora_get_hparams(self.h, C.byref(hp))
        self.hp = hp

    def close(self):
        if self.h:
            self.L.ora_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_act_round(self, mode):
        self.L.ora_set_act_round(self.h, mode)

    def set_threads(self, n):
        self.L.ora_set_threads(self.h, n)

    def token_str(self, i):
        return self.L.ora_token_to_str(self.h, i)

    def tokenize(self, text):
        buf = np.zeros(1024, np.int32)
        n = self.L.ora_tokenize(self.h, text.encode(), _ip(buf), 1024)
        return buf[:n].copy()

    def mel(self, pcm):
        pcm = np.ascontiguousarray(pcm, np.float32)
        n_len, n_org = C.c_int(), C.c_int()
        self.L.ora_mel(self.h, _fp(pcm), len(pcm), None, C.byref(n_len), C.byref(n_org))
        out = np.empty((self.hp.n_mels, n_len.value), np.float32)
        self.L.ora_mel(self.h, _fp(pcm), len(pcm), _fp(out), C.byref(n_len), C.byref(n_org))
        return out, n_org.value

    def encode(self, mel_window):
        mw = np.ascontiguousarray(mel_window, np.float32)
        assert mw.shape == (self.hp.n_mels, 2 * self.hp.n_audio_ctx)
        out = np.empty((self.hp.n_audio_ctx, self.hp.n_audio_state), np.float32)
        if self.L.ora_encode(self.h, _fp(mw), _fp(out)):
            raise RuntimeError(self.L.ora_last_error().decode())
        return out

    def encode_tap(self, which):
        out = np.empty((self.hp.n_audio_ctx, self.hp.n_audio_state), np.float32)
        if self.L.ora_encode_tap(self.h, which, _fp(out)) < 0:
            raise RuntimeError("bad tap")
        return out

    def decode(self, tokens, n_past=0, slot=0):
        tk = np.ascontiguousarray(tokens, np.int32)
        out = np.empty((len(tk), self.hp.n_vocab), np.float32)
        if self.L.ora_decode(self.h, slot, _ip(tk), len(tk), n_past, _fp(out)):
            raise RuntimeError(self.L.ora_last_error().decode())
        return out

    def default_params(self, strategy=0, **kw):
        p = self.L.ora_full_default_params(strategy)
        self._keep = []
        for k, v in kw.items():
            if isinstance(v, str):
                v = v.encode()
                self._keep.append(v)
            setattr(p, k, v)
        return p

    def process_logits(self, params, tokens_cur, has_ts, seek_delta, temperature, logits):
        n = self.hp.n_vocab
        tk = np.ascontiguousarray(tokens_cur, np.int32)
        lg = np.ascontiguousarray(logits, np.float32)
        lo, lp, pr = (np.empty(n, np.float32) for _ in range(3))
        self.L.ora_process_logits(self.h, C.byref(params), _ip(tk), len(tk), int(has_ts), seek_delta,
                                  temperature, _fp(lg), _fp(lo), _fp(lp), _fp(pr))
        return lo, lp, pr

    def full(self, pcm, params):
        pcm = np.ascontiguousarray(pcm, np.float32)
        r = C.c_void_p()
        rc = self.L.ora_full(self.h, C.byref(params), _fp(pcm), len(pcm), C.byref(r))
        if rc:
            raise RuntimeError("ora_full: " + self.L.ora_last_error().decode())
        L = self.L
        out = dict(segments=result_to_dict(L, r, "ora_", tok_to_dict),
                   lang_id=L.ora_result_lang_id(r), n_decode_steps=L.ora_result_n_decode_steps(r),
                   n_windows=L.ora_result_n_windows(r), ms_mel=L.ora_result_ms_mel(r),
                   ms_encode=L.ora_result_ms_encode(r), ms_decode=L.ora_result_ms_decode(r))
        L.ora_result_free(r)
        return out
