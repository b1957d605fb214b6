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


def make_script(sp, n_tokens, seed, n_vocab_text, end_cs=3000, final_pair=False):
    """A valid timestamp-token transcript of n_tokens sampled tokens ending in EOT.
    end_cs: last timestamp in centiseconds/2 units (1500 = 30.00 s)."""
    rng = np.random.default_rng(seed)
    beg = sp["beg"]
    script = []
    n_seg = max(1, n_tokens // 14)
    last = min(end_cs // 2, 1500)
    cuts = np.sort(rng.choice(np.arange(20, last - 20), size=n_seg - 1, replace=False)) if n_seg > 1 else []
    bounds = [0] + [int(c) for c in cuts] + [last]
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


class GgmlWriter:
    def __init__(self, path):
        self.f = open(path, "wb")

    def i32(self, *v):
        self.f.write(struct.pack("<%di" % len(v), *v))

    def tensor(self, name, arr, f16):
        arr = np.ascontiguousarray(arr)
        dims = list(arr.shape)
        nb = name.encode()
        ttype = 1 if f16 else 0
        self.i32(len(dims), len(nb), ttype)
        for d in reversed(dims):  # ggml order: ne[0] is the contiguous dim
            self.i32(d)
        self.f.write(nb)
        (arr.astype(np.float16) if f16 else arr.astype(np.float32)).tofile(self.f)

    def close(self):
        self.f.close()


def generate(path, size, seed=None, script_len=0, script_rms=0.5, qk_gain=4.0, ln_f_gain=None,
             script_end_cs=3000, w_std=0.02, emb_std=0.02, f32_all=False, verbose=False,
             script_final_pair=False):
    d, n_head, n_layer, n_mel, n_vocab = SIZES[size]
    if seed is None:
        seed = int.from_bytes(hashlib.sha256(size.encode()).digest()[:4], "little")
    rng = np.random.default_rng(seed)
    sp = special_tokens(n_vocab)
    wr = GgmlWriter(path)
    wr.f.write(struct.pack("<I", 0x67676D6C))
    ftype = 0 if f32_all else 1
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

    def normal(shape, std):
        return rng.standard_normal(size=shape, dtype=np.float32) * np.float32(std)

    def attn(prefix, cross=False):
        wr.tensor(prefix + ".query.weight", normal((d, d), sw * qk_gain), f16)
        wr.tensor(prefix + ".query.bias", normal((d,), 0.01), False)
        wr.tensor(prefix + ".key.weight", normal((d, d), sw * qk_gain), f16)
        wr.tensor(prefix + ".value.weight", normal((d, d), sw), f16)
        wr.tensor(prefix + ".value.bias", normal((d,), 0.01), False)
        wr.tensor(prefix + ".out.weight", normal((d, d), sw), f16)
        wr.tensor(prefix + ".out.bias", normal((d,), 0.01), False)

    def ln(prefix, gain=1.0):
        wr.tensor(prefix + ".weight", np.full((d,), gain, np.float32), False)
        wr.tensor(prefix + ".bias", np.zeros((d,), np.float32), False)

    def mlp(prefix):
        wr.tensor(prefix + ".0.weight", normal((4 * d, d), sw), f16)
        wr.tensor(prefix + ".0.bias", normal((4 * d,), 0.01), False)
        wr.tensor(prefix + ".2.weight", normal((d, 4 * d), sw), f16)
        wr.tensor(prefix + ".2.bias", normal((d,), 0.01), False)

    # ---- encoder
    wr.tensor("encoder.positional_embedding", sinusoids(N_AUDIO_CTX, d), False)
    # conv weights larger so that the stem output has O(1) scale
    wr.tensor("encoder.conv1.weight", normal((d, n_mel, 3), 1.0 / np.sqrt(3 * n_mel)), f16)
    wr.tensor("encoder.conv1.bias", normal((d, 1), 0.01), False)
    wr.tensor("encoder.conv2.weight", normal((d, d, 3), 1.0 / np.sqrt(3 * d)), f16)
    wr.tensor("encoder.conv2.bias", normal((d, 1), 0.01), False)
    for i in range(n_layer):
        p = "encoder.blocks.%d" % i
        ln(p + ".attn_ln")
        attn(p + ".attn")
        ln(p + ".mlp_ln")
        mlp(p + ".mlp")
    ln("encoder.ln_post")

    # ---- decoder
    tok_emb = normal((n_vocab, d), emb_std)
    pos_emb = normal((N_TEXT_CTX, d), 0.01)
    script = []
    if script_len > 0:
        script = make_script(sp, script_len, seed + 2, n_base - 1, end_cs=script_end_cs,
                             final_pair=script_final_pair)
        # first sampled position: 3 for multilingual ([sot, lang, transcribe]), 1 for .en ([sot])
        p0 = 3 if n_vocab >= 51865 else 1
        e16 = tok_emb.astype(np.float16).astype(np.float32)
        alpha = script_rms / emb_std
        for i, tok in enumerate(script):
            p = p0 - 1 + i  # input position whose output predicts script[i]
            if p >= N_TEXT_CTX:
                break
            pos_emb[p] += alpha * e16[tok]
        if ln_f_gain is None:
            # closed-form calibration (see module docstring): b = logit of the scripted token
            # at gain 1; choose gain so that it clears logsumexp of the rest by ~2.2 nats.
            sigma_x = np.sqrt(script_rms ** 2 + 0.04 * n_layer + 0.02 ** 2)
            b = script_rms * emb_std * d / sigma_x
            noise = emb_std * np.sqrt(d)
            best = 1.0
            for g in np.arange(0.5, 40.0, 0.05):
                if g * b >= np.log(n_vocab) + (g * noise) ** 2 / 2 + 2.2:
                    best = g
                    break
            ln_f_gain = float(best)
    if ln_f_gain is None:
        ln_f_gain = 1.0
    wr.tensor("decoder.positional_embedding", pos_emb, False)
    wr.tensor("decoder.token_embedding.weight", tok_emb, f16)
    del tok_emb
    for i in range(n_layer):
        p = "decoder.blocks.%d" % i
        ln(p + ".attn_ln")
        attn(p + ".attn")
        ln(p + ".cross_attn_ln")
        attn(p + ".cross_attn", cross=True)
        ln(p + ".mlp_ln")
        mlp(p + ".mlp")
    ln("decoder.ln", gain=ln_f_gain)
    wr.close()
    info = dict(path=path, size=size, seed=seed, d=d, n_head=n_head, n_layer=n_layer, n_mel=n_mel,
                n_vocab=n_vocab, script=script, ln_f_gain=ln_f_gain, special=sp)
    if verbose:
        print({k: v for k, v in info.items() if k != "script"}, "script_len", len(script))
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
    a = ap.parse_args()
    generate(a.out, a.size, a.seed, a.script, a.script_rms, a.qk_gain, a.ln_f_gain, a.script_end_cs,
             f32_all=a.f32, verbose=True)


if __name__ == "__main__":
    sys.exit(main())
