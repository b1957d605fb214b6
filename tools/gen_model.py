#!/usr/bin/env python3
"""Synthetic Whisper model generator: writes a legacy-ggml `.bin` file
(the format whisper.cpp v1.8.2 loads through whisper_init_from_file_with_params,
reference call site /root/reference/src/stt_engine.cpp:33; layout restated in
SURVEY.md Appendix A.2).

There are no model weights on disk and no network, so every model used by the
tests and the benchmark is produced here from a seed: real Slaney mel
filterbank, sinusoidal encoder positions, Gaussian weights stored f16, and a
vocabulary with the special tokens at their upstream ids.

`--script N`: makes the decoder *peaky* the way a trained model is. A random
decoder has flat logits (top-1/top-2 gap of ~0.2 sigma over 51865 tokens), so
two correct implementations with different rounding diverge after a few greedy
steps. With a script, decoder.positional_embedding[p] additionally carries
alpha * token_embedding[script[p+1]], so the greedy continuation at position p
is script[p+1] with probability ~0.9 and a margin far above bf16 noise, while
every kernel still contributes to the logit values, `p` and `plog`.
The script is a valid Whisper timestamp-token sequence
(<|0.00|> text.. <|t|><|t|> text.. <|t_end|> EOT).

`--keyed K` (with --script): a model that LISTENS. A scripted model emits its transcript whatever the
audio is, so token parity on it proves sequencing, not arithmetic. A keyed model has K alternative
tokens per text position and picks the one the AUDIO names, through the real path (mel -> conv stem ->
encoder -> ln_post -> cross-KV -> cross attention -> logits); only a few rows of otherwise random weight
matrices are designed:
  * encoder: K "band" channels carry the mean log-mel of K frequency bands through both convolutions
    (the layers never write to them: those rows of attn.out / mlp.2 are zero); a set of "position"
    channels keeps the sinusoidal positional embedding untouched the same way;
  * decoder: head 0 of the LAST layer's cross attention is an alignment head: its query reads sin/cos
    of the target time tau(p) from reserved channels of the decoder positional embedding, its key reads
    the encoder's position channels, so position p attends sharply to encoder frame tau(p); its value is
    the K band channels and its output projection writes beta * sum_k band_k * w_k into the stream;
  * the K alternatives of a text position share one base embedding and differ by delta * w_k, so the
    logit of alternative k exceeds the others' by g * beta * delta * (band_k - band_k') / sigma.
With synth_audio.keyed_clip (one tone burst per text token, in the band of the wanted alternative) the
transcript spells the clip's symbol sequence; every other head, layer and weight stays random and still
shapes the logit values, `p` and `plog`. Timestamps stay scripted.
"""
import argparse
import hashlib
import struct
import sys

import numpy as np

SIZES = {
    #            d   heads L   n_mel n_vocab
    "micro":    (128, 2, 2, 80, 51865),
    "tiny":     (384, 6, 4, 80, 51865),
    "tiny.en":  (384, 6, 4, 80, 51864),
    "base":     (512, 8, 6, 80, 51865),
    "small":    (768, 12, 12, 80, 51865),
    "medium":   (1024, 16, 24, 80, 51865),
    "large-v3": (1280, 20, 32, 128, 51866),
    "small-4l":  (768, 12, 4, 80, 51865),    # small / medium widths at reduced depth: config 3 / 4 property tests
    "medium-3l": (1024, 16, 3, 80, 51865),
    "large-v3-2l": (1280, 20, 2, 128, 51866),  # large-v3 widths, 2 layers: kernel profiling without the 3 GB file
    "large-v3-8l": (1280, 20, 8, 128, 51866),  # 8 layers: per-layer decoder-step time without the 3 GB file
}

N_AUDIO_CTX = 1500
N_TEXT_CTX = 448

# strings whisper.cpp's suppress_nst matches against the vocabulary
NON_SPEECH = ['"', "#", "(", ")", "*", "+", "/", ":", ";", "<", "=", ">", "@", "[", "\\", "]", "^",
              "_", "`", "{", "|", "}", "~", "「", "」", "『", "』", "<<", ">>", "<<<", ">>>", "--",
              "---", "-(", "-[", "('", '("', "((", "))", "(((", ")))", "[[", "]]", "{{", "}}", "♪♪",
              "♪♪♪", "♩", "♪", "♫", "♬", "♭", "♮", "♯"]


def special_tokens(n_vocab):
    """Upstream id layout."""
    multilingual = n_vocab >= 51865
    if not multilingual:
        return dict(eot=50256, sot=50257, translate=50357, transcribe=50358, solm=50359,
                    prev=50360, nosp=50361, not_=50362, beg=50363, n_langs=0)
    n_langs = n_vocab - 51765 - 1
    dt = n_langs - 98
    return dict(eot=50257, sot=50258, translate=50357 + dt, transcribe=50358 + dt,
                solm=50359 + dt, prev=50360 + dt, nosp=50361 + dt, not_=50362 + dt,
                beg=50363 + dt, n_langs=n_langs)


def hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    min_log_hz, min_log_mel, logstep = 1000.0, 15.0, 27.0 / np.log(6.4)
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) * logstep, mels)


def mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    min_log_hz, min_log_mel, logstep = 1000.0, 15.0, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f)


def mel_filterbank(n_mel, n_fft=400, sr=16000):
    """Slaney-normalised triangular filterbank, [n_mel][n_fft/2+1] f32 (librosa.filters.mel /
    OpenAI mel_filters.npz construction)."""
    n_bins = n_fft // 2 + 1
    fft_freqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_pts = np.linspace(hz_to_mel_slaney(0.0), hz_to_mel_slaney(sr / 2.0), n_mel + 2)
    hz_pts = mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(hz_pts)
    ramps = hz_pts[:, None] - fft_freqs[None, :]
    w = np.zeros((n_mel, n_bins), dtype=np.float64)
    for i in range(n_mel):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (hz_pts[2:n_mel + 2] - hz_pts[:n_mel])
    w *= enorm[:, None]
    return w.astype(np.float32)


def sinusoids(length, channels, max_timescale=10000.0):
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = np.exp(-inc * np.arange(channels // 2))
    t = np.arange(length)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)


def make_vocab(n_base, seed):
    """n_base byte strings: 256 single bytes, the non-speech symbols, then unique pseudo-words."""
    rng = np.random.default_rng(seed)
    toks, seen = [], set()
    for i in range(256):
        b = bytes([i])
        toks.append(b)
        seen.add(b)
    for s in NON_SPEECH + [" -", " '", " (", " [", " \"", " ♪"]:
        b = s.encode("utf-8")
        if b not in seen:
            toks.append(b)
            seen.add(b)
    letters = "etaoinshrdlucmfwypvbgkqjxz"
    probs = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.2, 2.4,
                      2.0, 1.9, 1.0, 1.5, 2.0, 0.8, 0.1, 0.15, 0.15, 0.07])
    probs = probs / probs.sum()
    while len(toks) < n_base:
        n = int(rng.integers(1, 8))
        w = "".join(rng.choice(list(letters), size=n, p=probs))
        if rng.random() < 0.6:
            w = " " + w
        if rng.random() < 0.08:
            w = w.title() if not w.startswith(" ") else " " + w[1:].title()
        b = w.encode()
        if b in seen:
            continue
        seen.add(b)
        toks.append(b)
    return toks[:n_base]


def make_script(sp, n_tokens, seed, n_vocab_text, end_cs=3000, final_pair=False, first_ts=0):
    """A valid timestamp-token transcript of n_tokens sampled tokens ending in EOT.
    end_cs: last timestamp in centiseconds/2 units (1500 = 30.00 s)."""
    rng = np.random.default_rng(seed)
    beg = sp["beg"]
    script = []
    n_seg = max(1, n_tokens // 14)
    last = min(end_cs // 2, 1500)
    cuts = np.sort(rng.choice(np.arange(20, last - 20), size=n_seg - 1, replace=False)) if n_seg > 1 else []
    bounds = [int(first_ts)] + [int(c) for c in cuts if int(c) > first_ts] + [last]
    n_seg = len(bounds) - 1
    bounds[-1] = min(bounds[-1], 1500)
    # budget: per segment 2 timestamps + text; final EOT
    n_text_total = n_tokens - 1 - 2 * n_seg
    per = [n_text_total // n_seg] * n_seg
    for i in range(n_text_total - sum(per)):
        per[i] += 1
    for s in range(n_seg):
        script.append(beg + bounds[s])
        # ordinary text tokens (avoid the single-byte/non-speech region and specials)
        script.extend(int(t) for t in rng.integers(400, n_vocab_text, size=per[s]))
        script.append(beg + bounds[s + 1])
    if final_pair:  # transcript ends on a timestamp PAIR: upstream then seeks to it (second window)
        script.append(script[-1])
    script.append(sp["eot"])
    return script


KEYED_TONES_HZ = [410.0, 1020.0, 2200.0, 4400.0, 650.0, 1500.0, 3100.0, 6000.0]
KEYED_BAND_REL = 0.09   # band k = the mel bins whose centre lies within +-9 % of tone k (clips jitter by +-3 %)


def keyed_tone_hz(k):
    return KEYED_TONES_HZ[k]


def probe_band_levels(fb, band_bins, tone_hz, amp, noise, sr=16000):
    """Normalised log-mel ((log10 P + 4) / 4, whisper's scaling without the max - 8 clamp) averaged over
    band_bins for a steady tone of `amp` in white noise `noise`, and for the noise alone."""
    rng = np.random.default_rng(12345)
    n = sr
    t = np.arange(n) / sr
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 400)
    out = []
    for a in (amp, 0.0):
        x = a * np.sin(2 * np.pi * tone_hz * t) + rng.normal(0.0, noise, n)
        fr = np.stack([x[i:i + 400] * win for i in range(0, n - 400, 160)])
        pw = np.abs(np.fft.rfft(fr, axis=1)) ** 2
        mel = np.log10(np.maximum(pw @ fb.T.astype(np.float64), 1e-10))
        out.append(float(np.median(((mel + 4.0) / 4.0)[:, band_bins].mean(axis=1))))
    return out[0], out[1]


def mel_centers_hz(n_mel, sr=16000):
    mel_pts = np.linspace(hz_to_mel_slaney(0.0), hz_to_mel_slaney(sr / 2.0), n_mel + 2)
    return mel_to_hz_slaney(mel_pts)[1:-1]


def make_keyed_script(sp, n_tokens, seed, n_vocab_text, K, min_frames=16):
    """A timestamped transcript whose text positions have K alternatives and a slot of encoder frames
    each. Returns (script with alternative 0, variants [n_text][K], slots [(f0, f1)] in 20 ms frames,
    text_index: script index of every text token)."""
    rng = np.random.default_rng(seed)
    beg = sp["beg"]
    n_seg = max(1, n_tokens // 14)
    n_text_total = n_tokens - 1 - 2 * n_seg
    assert n_text_total * min_frames <= 1500, "keyed script: %d text tokens do not fit 30 s" % n_text_total
    per = [n_text_total // n_seg] * n_seg
    for i in range(n_text_total - sum(per)):
        per[i] += 1
    # segment lengths proportional to their token counts, over the whole window
    edges = np.round(np.cumsum([0] + per) * (1500.0 / n_text_total)).astype(int)
    toks = rng.choice(np.arange(400, n_vocab_text), size=n_text_total * K, replace=False).reshape(n_text_total, K)
    script, slots, text_index = [], [], []
    it = 0
    for sgi in range(n_seg):
        script.append(beg + int(edges[sgi]))
        f = np.linspace(edges[sgi], edges[sgi + 1], per[sgi] + 1)
        for j in range(per[sgi]):
            text_index.append(len(script))
            script.append(int(toks[it, 0]))
            slots.append((float(f[j]), float(f[j + 1])))
            it += 1
        script.append(beg + int(edges[sgi + 1]))
    script.append(sp["eot"])
    return script, toks.astype(int).tolist(), slots, text_index


def keyed_expected_tokens(info, symbols):
    """The sampled tokens whisper_full keeps for a clip whose text slot i carries symbol symbols[i]
    (the second timestamp of a pair and EOT are dropped, as in the result segments)."""
    k = info["keyed"]
    script = list(info["script"])
    for i, si in enumerate(k["text_index"]):
        script[si] = k["variants"][i][symbols[i]]
    beg = info["special"]["beg"]
    return [t for i, t in enumerate(script[:-1]) if not (i > 0 and t >= beg and script[i - 1] == t)]


def keyed_sure_prefix(info, n_samples, guard=3200):
    """How many of keyed_expected_tokens() a clip cut to n_samples (16 kHz) still decides: up to the first text
    token whose tone slot is not wholly inside the audio (minus a 0.2 s guard) or the first timestamp past its end."""
    k = info["keyed"]
    script, beg = info["script"], info["special"]["beg"]
    slot_of = {si: k["slots"][i] for i, si in enumerate(k["text_index"])}
    kept = 0
    for i, t in enumerate(script[:-1]):
        if i > 0 and t >= beg and script[i - 1] == t:
            continue
        if i in slot_of:
            if slot_of[i][1] * 320 > n_samples - guard:
                return kept
        elif t >= beg and (t - beg) * 320 > n_samples:
            return kept
        kept += 1
    return kept


QUANT_FTYPE = {"q4_0": 2, "q4_1": 3, "q8_0": 7, "q5_0": 8, "q5_1": 9}   # ggml_ftype (file header)
QUANT_TTYPE = {"q4_0": 2, "q4_1": 3, "q5_0": 6, "q5_1": 7, "q8_0": 8}   # ggml_type (tensor header)


class GgmlWriter:
    def __init__(self, path, quant=None, twin=False):
        """quant: one of QUANT_TTYPE - 2-D `.weight` matrices are stored block-quantised, as whisper.cpp's
        quantize tool does (through the `gguf` package's reference quantisers). twin: store the DEQUANTISED
        values as f32 instead - a file that must load to the very same bf16 weights as the quantised one."""
        self.f = open(path, "wb")
        self.quant, self.twin = quant, twin

    def i32(self, *v):
        self.f.write(struct.pack("<%di" % len(v), *v))

    def tensor(self, name, arr, f16):
        arr = np.ascontiguousarray(arr)
        dims = list(arr.shape)
        nb = name.encode()
        ttype = 1 if f16 else 0
        data = None
        if self.quant and len(dims) == 2 and name.endswith(".weight") and dims[1] % 32 == 0 and f16:
            from gguf import GGMLQuantizationType, quants
            qt = GGMLQuantizationType(QUANT_TTYPE[self.quant])
            q = quants.quantize(arr.astype(np.float32), qt)
            if self.twin:
                ttype, data = 0, quants.dequantize(q, qt).astype(np.float32)
            else:
                ttype, data = QUANT_TTYPE[self.quant], q
        self.i32(len(dims), len(nb), ttype)
        for d in reversed(dims):  # ggml order: ne[0] is the contiguous dim
            self.i32(d)
        self.f.write(nb)
        if data is not None:
            np.ascontiguousarray(data).tofile(self.f)
        else:
            (arr.astype(np.float16) if f16 else arr.astype(np.float32)).tofile(self.f)

    def close(self):
        self.f.close()


def generate(path, size, seed=None, script_len=0, script_rms=0.5, qk_gain=4.0, ln_f_gain=None,
             script_end_cs=3000, w_std=0.02, emb_std=0.02, f32_all=False, verbose=False,
             script_final_pair=False, script_first_ts=0, keyed=0, keyed_beta=1.2, keyed_delta=0.6, keyed_attn=1.5,
             keyed_gain=0.7, quant=None, quant_twin=False):
    d, n_head, n_layer, n_mel, n_vocab = SIZES[size]
    if seed is None:
        seed = int.from_bytes(hashlib.sha256(size.encode()).digest()[:4], "little")
    rng = np.random.default_rng(seed)
    sp = special_tokens(n_vocab)
    wr = GgmlWriter(path, quant=quant, twin=quant_twin)
    wr.f.write(struct.pack("<I", 0x67676D6C))
    ftype = 0 if f32_all else 1
    if quant and not quant_twin:
        ftype = 2000 + QUANT_FTYPE[quant]  # quantisation version 2 in the thousands, as ggml writes it
    wr.i32(n_vocab, N_AUDIO_CTX, d, n_head, n_layer, N_TEXT_CTX, d, n_head, n_layer, n_mel, ftype)
    fb = mel_filterbank(n_mel)
    wr.i32(n_mel, 201)
    fb.tofile(wr.f)
    n_base = sp["eot"] + (1 if n_vocab >= 51865 else 0)  # upstream files list the tokenizer vocab only
    n_base = 50257 if n_vocab >= 51865 else 50256
    vocab = make_vocab(n_base, seed + 1)
    wr.i32(len(vocab))
    for t in vocab:
        wr.f.write(struct.pack("<I", len(t)))
        wr.f.write(t)

    sw = w_std * np.sqrt(384.0 / d)
    f16 = not f32_all

    # ---- keyed ("listening") model: reserved channels and designed rows (module docstring)
    K = int(keyed)
    kp = None
    if K:
        assert script_len > 0 and 2 <= K <= len(KEYED_TONES_HZ) and d // n_head == 64
        half = d // 2
        inc = np.log(10000.0) / (half - 1)
        omega = np.exp(-inc * np.arange(half))
        # alignment kernel sum_j cos(omega_j dt): half width ~ 4 / omega_max frames; text slots are
        # 1500 / n_text frames wide (43 for a 40-token script, 17.6 for a 100-token one)
        n_text_est = script_len - 1 - 2 * max(1, script_len // 14)
        om_max = 0.25 if 1500.0 / max(1, n_text_est) >= 30 else 0.55
        cand = np.flatnonzero((omega >= 0.003) & (omega <= om_max))
        J = min(32, max(12, d // 16), len(cand))  # sin/cos pairs of the alignment head (2J <= 64 head dims)
        js = cand[np.round(np.linspace(0, len(cand) - 1, J)).astype(int)]
        enc_pos_ch = np.concatenate([js, half + js])            # sin | cos columns of the sinusoid table
        enc_band_ch = np.array([half - 1 - k for k in range(K)])  # slowest sin columns (overwritten with 0)
        enc_zero_ch = half - 1 - K                              # stays 0: LayerNorm's (0 - mean) / std = the offset
        dec_q_ch = np.arange(d - 2 * J, d)                      # decoder channels that carry sin | cos of tau(p)
        dec_zero_ch = d - 2 * J - 1                             # to subtract from every other reserved channel
        centers = mel_centers_hz(n_mel)
        band_bins = [np.flatnonzero(np.abs(centers / f - 1.0) <= KEYED_BAND_REL) for f in KEYED_TONES_HZ[:K]]
        assert all(len(b) > 0 for b in band_bins)
        # per band: level of a synth_audio.keyed_clip tone (amp 0.34 = middle of its range, noise 0.003)
        # and of the noise floor, so that every band channel reads ~0 when idle and ~1.5 when active
        band_levels = [probe_band_levels(fb, band_bins[k], KEYED_TONES_HZ[k], 0.34, 0.003) for k in range(K)]
        kp = dict(J=J, omega=omega[js], enc_pos_ch=enc_pos_ch, enc_band_ch=enc_band_ch, dec_q_ch=dec_q_ch,
                  enc_zero_ch=enc_zero_ch, dec_zero_ch=dec_zero_ch,
                  enc_reserved=np.concatenate([enc_pos_ch, enc_band_ch, [enc_zero_ch]]),
                  dec_reserved=np.concatenate([dec_q_ch, [dec_zero_ch]]))

    def normal(shape, std):
        return rng.standard_normal(size=shape, dtype=np.float32) * np.float32(std)

    def zero_rows(w, rows):
        if rows is not None:
            w[rows] = 0.0
        return w

    def attn(prefix, out_zero=None, edit=None):
        t = dict(qw=normal((d, d), sw * qk_gain), qb=normal((d,), 0.01), kw=normal((d, d), sw * qk_gain),
                 vw=normal((d, d), sw), vb=normal((d,), 0.01), ow=normal((d, d), sw), ob=normal((d,), 0.01))
        zero_rows(t["ow"], out_zero)
        zero_rows(t["ob"], out_zero)
        if edit:
            edit(t)
        wr.tensor(prefix + ".query.weight", t["qw"], f16)
        wr.tensor(prefix + ".query.bias", t["qb"], False)
        wr.tensor(prefix + ".key.weight", t["kw"], f16)
        wr.tensor(prefix + ".value.weight", t["vw"], f16)
        wr.tensor(prefix + ".value.bias", t["vb"], False)
        wr.tensor(prefix + ".out.weight", t["ow"], f16)
        wr.tensor(prefix + ".out.bias", t["ob"], False)

    def ln(prefix, gain=1.0):
        wr.tensor(prefix + ".weight", np.full((d,), gain, np.float32), False)
        wr.tensor(prefix + ".bias", np.zeros((d,), np.float32), False)

    def mlp(prefix, out_zero=None):
        wr.tensor(prefix + ".0.weight", normal((4 * d, d), sw), f16)
        wr.tensor(prefix + ".0.bias", normal((4 * d,), 0.01), False)
        wr.tensor(prefix + ".2.weight", zero_rows(normal((d, 4 * d), sw), out_zero), f16)
        wr.tensor(prefix + ".2.bias", zero_rows(normal((d,), 0.01), out_zero), False)

    # ---- encoder
    enc_pos = sinusoids(N_AUDIO_CTX, d)
    # conv weights larger so that the stem output has O(1) scale
    c1w, c1b = normal((d, n_mel, 3), 1.0 / np.sqrt(3 * n_mel)), normal((d, 1), 0.01)
    c2w, c2b = normal((d, d, 3), 1.0 / np.sqrt(3 * d)), normal((d, 1), 0.01)
    if K:
        enc_pos[:, kp["enc_band_ch"]] = 0.0
        enc_pos[:, kp["enc_zero_ch"]] = 0.0
        for k in range(K):  # band channel k = mean log-mel of band k over 3 frames, passed through conv2 as is
            ch = kp["enc_band_ch"][k]
            c1w[ch] = 0.0
            act, idle = band_levels[k]
            gain = 1.5 / (act - idle)
            c1w[ch, band_bins[k], :] = gain / (3 * len(band_bins[k]))
            c1b[ch] = -gain * idle
            c2w[ch] = 0.0
            c2w[ch, ch, :] = 1.0 / 3.0
            c2b[ch] = 0.0
        for chs in (kp["enc_pos_ch"], [kp["enc_zero_ch"]]):  # position channels: GELU(0) = 0, + the sinusoids
            c2w[chs] = 0.0
            c2b[chs] = 0.0
    wr.tensor("encoder.positional_embedding", enc_pos, False)
    wr.tensor("encoder.conv1.weight", c1w, f16)
    wr.tensor("encoder.conv1.bias", c1b, False)
    wr.tensor("encoder.conv2.weight", c2w, f16)
    wr.tensor("encoder.conv2.bias", c2b, False)
    del c1w, c2w
    enc_keep = kp["enc_reserved"] if K else None
    for i in range(n_layer):
        p = "encoder.blocks.%d" % i
        ln(p + ".attn_ln")
        attn(p + ".attn", out_zero=enc_keep)
        ln(p + ".mlp_ln")
        mlp(p + ".mlp", out_zero=enc_keep)
    ln("encoder.ln_post")

    # ---- decoder
    tok_emb = normal((n_vocab, d), emb_std)
    pos_emb = normal((N_TEXT_CTX, d), 0.01)
    script = []
    keyed_info = None
    extra_var = 0.0  # mean square per element the designed terms add to the final hidden state
    if script_len > 0:
        # first sampled position: 3 for multilingual ([sot, lang, transcribe]), 1 for .en ([sot])
        p0 = 3 if n_vocab >= 51865 else 1
        alpha = script_rms / emb_std
        if K:
            script, variants, slots, text_index = make_keyed_script(sp, script_len, seed + 2, n_base - 1, K)
            J, qch = kp["J"], kp["dec_q_ch"]
            rch = kp["dec_reserved"]
            wrng = np.random.default_rng(seed + 3)
            wdir = wrng.standard_normal((K, d))
            wdir[:, rch] = 0.0
            wdir = np.linalg.qr(wdir.T)[0].T.astype(np.float32)   # K orthonormal directions, zero on the q channels
            kp["wdir"] = wdir
            tok_emb[:, rch] = 0.0
            pos_emb[:, rch] = 0.0
            delta = keyed_delta * emb_std * np.sqrt(d)            # |delta w_k| = keyed_delta * |e|
            e16 = tok_emb.astype(np.float16).astype(np.float32)
            ti = 0
            for i, tok in enumerate(script):
                p = p0 - 1 + i
                if p >= N_TEXT_CTX:
                    break
                pos_emb[p] += alpha * e16[tok]                    # the base embedding (before delta w_k is added)
                if ti < len(text_index) and text_index[ti] == i:
                    tau = 0.5 * (slots[ti][0] + slots[ti][1])
                    pos_emb[p, qch[:J]] = np.sin(kp["omega"] * tau)
                    pos_emb[p, qch[J:]] = np.cos(kp["omega"] * tau)
                    ti += 1
            for i, alts in enumerate(variants):                   # alternatives share the base, differ by delta w_k
                base = tok_emb[alts[0]].copy()
                for k, t in enumerate(alts):
                    tok_emb[t] = base + delta * wdir[k]
            keyed_info = dict(K=K, variants=variants, slots=slots, text_index=text_index,
                              tones_hz=[keyed_tone_hz(k) for k in range(K)])
            # keyed_beta is relative: the band term written into the stream has about keyed_beta times the
            # norm of the scripted term. An active band reads ~1.29 / sigma_t after ln_post (idle: 0), where
            # sigma_t^2 ~ 0.2 + 0.036 L is the encoder stream's variance (measured on the oracle, L = 2 .. 24);
            # the mean-free value vector has sqrt(3/4) of that norm
            sigma_t = np.sqrt(0.2 + 0.036 * n_layer)
            kp["beta"] = keyed_beta * script_rms * np.sqrt(d) / (1.29 / sigma_t * np.sqrt(1.0 - 1.0 / K))
            extra_var = (J * 1.0) / d + (keyed_beta * script_rms) ** 2
        else:
            script = make_script(sp, script_len, seed + 2, n_base - 1, end_cs=script_end_cs,
                                 final_pair=script_final_pair, first_ts=script_first_ts)
            e16 = tok_emb.astype(np.float16).astype(np.float32)
            for i, tok in enumerate(script):
                p = p0 - 1 + i  # input position whose output predicts script[i]
                if p >= N_TEXT_CTX:
                    break
                pos_emb[p] += alpha * e16[tok]
        if ln_f_gain is None:
            # closed-form calibration (see module docstring): b = logit of the scripted token
            # at gain 1; choose gain so that it clears logsumexp of the rest by ~2.2 nats.
            sigma_x = np.sqrt(script_rms ** 2 + 0.04 * n_layer + 0.02 ** 2 + extra_var)
            b = script_rms * emb_std * d / sigma_x
            noise = emb_std * np.sqrt(d)
            best = 1.0
            for g in np.arange(0.5, 40.0, 0.05):
                if g * b >= np.log(n_vocab) + (g * noise) ** 2 / 2 + 2.2:
                    best = g
                    break
            ln_f_gain = float(best) * (keyed_gain if K else 1.0)  # keyed: the band term adds to the winner's margin
    if ln_f_gain is None:
        ln_f_gain = 1.0
    wr.tensor("decoder.positional_embedding", pos_emb, False)
    wr.tensor("decoder.token_embedding.weight", tok_emb, f16)
    del tok_emb
    dec_keep = kp["dec_reserved"] if K else None

    def alignment_head(t):
        # head 0 (rows / columns 0..63) of the last layer's cross attention
        # score(p, t) = a^2 / 8 * sum_r q_r k_r / (sigma' sigma_t) = s * sum_j cos(omega_j (t - tau_p)): the sharpness s
        # (keyed_attn) is kept the same for every depth by scaling a with the two LayerNorm divisors
        sigma_q = np.sqrt(script_rms ** 2 + 0.04 * n_layer + extra_var)
        J, a = kp["J"], np.float32(np.sqrt(8.0 * keyed_attn * sigma_q * np.sqrt(0.2 + 0.036 * n_layer)))
        for name in ("qw", "kw", "vw"):
            t[name][:64] = 0.0
        t["qb"][:64] = 0.0
        t["vb"][:64] = 0.0
        # every read subtracts the always-zero channel: LayerNorm maps a channel x to (x - mean) / std, so the
        # difference is x / std - the data-dependent offset is gone (q is exactly 0 where pos_emb holds no time)
        for r in range(2 * J):
            t["qw"][r, kp["dec_q_ch"][r]] = a
            t["qw"][r, kp["dec_zero_ch"]] = -a
            t["kw"][r, kp["enc_pos_ch"][r]] = a
            t["kw"][r, kp["enc_zero_ch"]] = -a
        # value k = band_k - mean of the bands: zero when the head looks everywhere at once (timestamp positions,
        # q = 0) and clips use the bands evenly; the coefficients sum to 0, so LayerNorm's offset cancels here too
        for k in range(K):
            for k2 in range(K):
                t["vw"][k, kp["enc_band_ch"][k2]] = (1.0 if k == k2 else 0.0) - 1.0 / K
        t["ow"][:, :64] = 0.0
        for k in range(K):
            t["ow"][:, k] = kp["beta"] * kp["wdir"][k]

    for i in range(n_layer):
        p = "decoder.blocks.%d" % i
        ln(p + ".attn_ln")
        attn(p + ".attn", out_zero=dec_keep)
        ln(p + ".cross_attn_ln")
        attn(p + ".cross_attn", out_zero=dec_keep, edit=alignment_head if (K and i == n_layer - 1) else None)
        ln(p + ".mlp_ln")
        mlp(p + ".mlp", out_zero=dec_keep)
    ln("decoder.ln", gain=ln_f_gain)
    wr.close()
    info = dict(path=path, size=size, seed=seed, d=d, n_head=n_head, n_layer=n_layer, n_mel=n_mel,
                n_vocab=n_vocab, script=script, ln_f_gain=ln_f_gain, special=sp)
    if keyed_info:
        info["keyed"] = keyed_info
    if verbose:
        print({k: v for k, v in info.items() if k not in ("script", "keyed")}, "script_len", len(script))
    return info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="tiny", choices=sorted(SIZES))
    ap.add_argument("--out", required=True)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--script", type=int, default=0, help="scripted (peaky) transcript length in tokens")
    ap.add_argument("--script-rms", type=float, default=0.5)
    ap.add_argument("--script-end-cs", type=int, default=3000)
    ap.add_argument("--qk-gain", type=float, default=4.0)
    ap.add_argument("--ln-f-gain", type=float, default=None)
    ap.add_argument("--f32", action="store_true")
    ap.add_argument("--keyed", type=int, default=0, help="alternatives per text position chosen by the audio (see docstring)")
    a = ap.parse_args()
    generate(a.out, a.size, a.seed, a.script, a.script_rms, a.qk_gain, a.ln_f_gain, a.script_end_cs,
             f32_all=a.f32, verbose=True, keyed=a.keyed)


if __name__ == "__main__":
    sys.exit(main())
