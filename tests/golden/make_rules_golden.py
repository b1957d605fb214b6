#!/usr/bin/env python3
"""Pins the oracle's LOGIT RULES and greedy SEQUENCING against an independent implementation and writes
tests/golden/rules_hf.npz.

The rules live in un-vendored whisper.cpp v1.8.2 (`whisper_process_logits`, called per decoder step from
`whisper_full_with_state`, reference call site /root/reference/src/stt_engine.cpp:245); the reference holds
no tests for them. The independent implementation available offline is HuggingFace
`transformers`: `SuppressTokensLogitsProcessor`, `SuppressTokensAtBeginLogitsProcessor` and
`WhisperTimeStampLogitsProcessor` (generation/logits_process.py:1812-2046, a port of OpenAI's
decoding.py rules that whisper.cpp also ports), and `WhisperForConditionalGeneration.generate`.

What is produced (all driven by HF code only; the oracle is compared against it in
tests/test_oracle_vs_hf_rules.py, which also re-runs HF live when transformers is importable):
  * for N seeded (logits, token history) cases: the -inf mask HF's three processors leave, as packed bits;
  * HF greedy `generate(return_timestamps=True)` token sequences on the keyed micro model (tools/gen_model.py
    --keyed: the audio chooses the text tokens) for 3 clips that spell different symbol sequences.

Intentional, NAMED differences between whisper.cpp (as restated by the oracle) and HF/OpenAI:
  D1 initial_text_not_forced : OpenAI/HF suppress every non-timestamp token at the first sampled position;
       whisper.cpp does not (a timestamp is forced only through the timestamp-mass rule).
  D2 equal_timestamp_allowed : after text that follows a timestamp, HF forbids timestamps <= the last one
       (`timestamp_last = timestamps[-1] + 1`); whisper.cpp forbids only timestamps < seek_delta/2, and
       tracks it only for tokens > <|0.00|> (has_ts), so repeating the last timestamp stays legal.
The case generator neutralises D1 / D2 in the *input* (initial cases have text logits far below the
timestamps, so that the mass rule forces a timestamp on both sides; the one id D2 concerns is pre-masked),
so the committed masks must match bit for bit; two extra cases exercise D1 and D2 on their own.
Run in the dev container:   python tests/golden/make_rules_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tools import gen_model, ggml_io, synth_audio  # noqa: E402

MODEL_ARGS = dict(size="micro", seed=1234, script_len=40, keyed=4)  # a model that listens: every clip, its own tokens
N_CASES = 96
NON_SPEECH = gen_model.NON_SPEECH


def static_suppress_ids(vocab, sp, suppress_nst=True, tdrz=False):
    """The ids whisper.cpp masks at every step, computed from the vocabulary alone (the role OpenAI's
    `suppress_tokens` list plays for HF): <|notimestamps|> is left to the timestamp processor."""
    ids = {sp["sot"], sp["nosp"], sp["translate"], sp["transcribe"], sp["prev"]}
    if not tdrz:
        ids.add(sp["solm"])
    ids.update(range(sp["sot"] + 1, sp["sot"] + 1 + sp["n_langs"]))
    if suppress_nst:
        index = {}
        for i, t in enumerate(vocab):
            index.setdefault(t, i)
        for s in NON_SPEECH:
            for v in (s, " " + s):
                if v.encode("utf-8") in index:
                    ids.add(index[v.encode("utf-8")])
        for v in (b" -", b" '"):
            if v in index:
                ids.add(index[v])
    return sorted(ids)


def space_id(vocab):
    return vocab.index(b" ")


def random_history(rng, sp, n_text_vocab):
    """A grammatical sampled-token history (after the prompt): <|t0|> text.. <|t|><|t|> text.. [<|t|>]"""
    beg = sp["beg"]
    kind = int(rng.integers(0, 6))
    if kind == 0:
        return []
    h = [beg + int(rng.integers(0, 3))]
    t = h[0] - beg
    n_seg = int(rng.integers(0, 3))
    for _ in range(n_seg):
        h += [int(x) for x in rng.integers(400, n_text_vocab, size=int(rng.integers(1, 5)))]
        t = min(1500, t + int(rng.integers(1, 200)))
        h += [beg + t, beg + t]
    if kind == 1:
        return h                                   # ends on a pair (or the lone initial timestamp)
    h += [int(x) for x in rng.integers(400, n_text_vocab, size=int(rng.integers(1, 5)))]
    if kind in (2, 3):
        return h                                   # ends on text
    t = min(1500, t + int(rng.integers(0, 200)))
    return h + [beg + t]                           # text then ONE timestamp: the pair must be closed


def make_case(rng, sp, n_vocab, n_text_vocab):
    hist = random_history(rng, sp, n_text_vocab)
    logits = rng.normal(0, 2.0, n_vocab).astype(np.float32)
    beg = sp["beg"]
    if not hist:
        logits[:beg] -= 40.0                       # D1 neutralised: the timestamp mass wins on both sides
    ts = [t for t in hist if t >= beg]
    if ts and not (len(hist) >= 2 and hist[-1] >= beg and hist[-2] < beg):
        logits[ts[-1]] = -np.inf                   # D2 neutralised: the one id the two rules disagree on
    if rng.random() < 0.3:
        logits[beg:] += 6.0                        # timestamp mass above any text token
    return hist, logits


def hf_processors(sp, vocab, n_prompt, suppress_nst=True):
    from transformers import GenerationConfig
    from transformers.generation.logits_process import (SuppressTokensAtBeginLogitsProcessor,
                                                        SuppressTokensLogitsProcessor,
                                                        WhisperTimeStampLogitsProcessor)
    gc = GenerationConfig(eos_token_id=sp["eot"], bos_token_id=sp["eot"])
    gc.no_timestamps_token_id = sp["not_"]
    gc.max_initial_timestamp_index = 50            # 1.0 s / 0.02 s (whisper.cpp: max_initial_ts = 1.0)
    return [SuppressTokensLogitsProcessor(static_suppress_ids(vocab, sp, suppress_nst)),
            SuppressTokensAtBeginLogitsProcessor([space_id(vocab), sp["eot"]], begin_index=n_prompt),
            WhisperTimeStampLogitsProcessor(gc, begin_index=n_prompt)]


def hf_mask(procs, prompt, hist, logits):
    ids = torch.tensor([list(prompt) + list(hist)], dtype=torch.long)
    s = torch.from_numpy(logits[None].copy())
    for p in procs:
        s = p(ids, s)
    return torch.isinf(s[0]).numpy() & (s[0].numpy() < 0), s[0].numpy()


def cases(sp, n_vocab, n_text_vocab, seed=20261018, n=N_CASES):
    rng = np.random.default_rng(seed)
    return [make_case(rng, sp, n_vocab, n_text_vocab) for _ in range(n)]


def hf_generate(m, sp, mel_window, max_new=60):
    """Greedy HF generation with timestamps over one 30 s window of log-mel [n_mel][3000]."""
    gc = m.generation_config
    gc.eos_token_id = sp["eot"]
    gc.pad_token_id = sp["eot"]
    gc.decoder_start_token_id = sp["sot"]
    gc.no_timestamps_token_id = sp["not_"]
    gc.max_initial_timestamp_index = 50
    gc.lang_to_id = {"<|en|>": sp["sot"] + 1}
    gc.task_to_id = {"transcribe": sp["transcribe"], "translate": sp["translate"]}
    gc.is_multilingual = True
    gc.prev_sot_token_id = sp["prev"]
    gc.return_timestamps = True
    gc.suppress_tokens = None
    gc.begin_suppress_tokens = None
    out = m.generate(input_features=torch.from_numpy(mel_window[None]), return_timestamps=True, language="en",
                     task="transcribe", do_sample=False, num_beams=1, max_new_tokens=max_new,
                     suppress_tokens=m._sw_suppress, begin_suppress_tokens=m._sw_begin_suppress)
    seq = out[0].tolist() if not isinstance(out, dict) else out["sequences"][0].tolist()
    return seq


def golden_clip(info, i):
    k = info["keyed"]
    return synth_audio.keyed_clip(k, synth_audio.keyed_symbols(k, 900 + i), seed=900 + i)


def main():
    from make_golden import hf_model
    from oracle import ora
    out_dir = os.path.dirname(os.path.abspath(__file__))
    path = "/tmp/sw_golden_micro.bin"
    info = gen_model.generate(path, **MODEL_ARGS)
    sp = info["special"]
    hp, _, vocab, _ = ggml_io.read_ggml(path)
    n_vocab = hp["n_vocab"]
    prompt = [sp["sot"], sp["sot"] + 1, sp["transcribe"]]
    procs = hf_processors(sp, vocab, len(prompt))
    cs = cases(sp, n_vocab, len(vocab) - 1)
    masks = np.stack([hf_mask(procs, prompt, h, lg)[0] for h, lg in cs])
    print("rule cases: %d, mean suppressed fraction %.3f" % (len(cs), masks.mean()))

    torch.set_grad_enabled(False)
    m, _, _ = hf_model(path)
    m._sw_suppress = static_suppress_ids(vocab, sp)
    m._sw_begin_suppress = [space_id(vocab), sp["eot"]]
    o = ora.Oracle(path, weight_round=False, act_round=ora.ACT_F32, gelu_erf=True)
    seqs = []
    for i in range(3):
        mel, _ = o.mel(synth_audio.to_f32(golden_clip(info, i)))
        seq = hf_generate(m, sp, mel[:, :3000])
        want = gen_model.keyed_expected_tokens(info, synth_audio.keyed_symbols(info["keyed"], 900 + i))
        kept = [t for j, t in enumerate(seq) if t != sp["eot"] and not (j > 0 and t >= sp["beg"] and seq[j - 1] == t)]
        assert kept == want, "HF does not read clip %d's symbols back" % i
        print("hf generate clip %d: %d tokens, head %s" % (i, len(seq), seq[:8]))
        seqs.append(np.array(seq, np.int32))
    n = max(len(s) for s in seqs)
    seq_arr = np.full((3, n), -1, np.int32)
    for i, s in enumerate(seqs):
        seq_arr[i, : len(s)] = s
    np.savez_compressed(os.path.join(out_dir, "rules_hf.npz"), model_args=np.array(repr(MODEL_ARGS)),
                        n_cases=np.array(len(cs)), masks=np.packbits(masks, axis=1), hf_sequences=seq_arr,
                        prompt=np.array(prompt, np.int32))
    print("wrote", os.path.join(out_dir, "rules_hf.npz"))


if __name__ == "__main__":
    main()
