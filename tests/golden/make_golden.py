#!/usr/bin/env python3
"""Pins the CPU oracle against an INDEPENDENT implementation and writes the golden fixtures.

The reference's own arithmetic (whisper.cpp v1.8.2) is not in /root/reference and the reference
holds no golden vectors, so parity is anchored on:
  * HuggingFace `transformers` Whisper (PyTorch fp32, CPU) fed the SAME synthetic ggml weights:
    encoder output and teacher-forced decoder logits (HF uses erf-GELU: the oracle's gelu_erf
    switch is flipped for this comparison only; its activation rounding is switched off);
  * HF WhisperFeatureExtractor for the log-mel (frames whose STFT window does not reach the end
    of the clip: upstream zero-pads there, HF reflects).
Run in the dev container (needs torch + transformers):   python tests/golden/make_golden.py
Output: tests/golden/micro_hf.npz  (self-generated, not upstream-verified).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ora  # noqa: E402
from tools import gen_model, ggml_io, synth_audio  # noqa: E402

MODEL_ARGS = dict(size="micro", seed=1234, script_len=40)


def hf_model(path):
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    hp, filters, vocab, T = ggml_io.read_ggml(path)
    cfg = WhisperConfig(
        vocab_size=hp["n_vocab"], num_mel_bins=hp["n_mels"], d_model=hp["n_audio_state"],
        encoder_layers=hp["n_audio_layer"], encoder_attention_heads=hp["n_audio_head"],
        decoder_layers=hp["n_text_layer"], decoder_attention_heads=hp["n_text_head"],
        encoder_ffn_dim=4 * hp["n_audio_state"], decoder_ffn_dim=4 * hp["n_text_state"],
        max_source_positions=hp["n_audio_ctx"], max_target_positions=hp["n_text_ctx"],
        activation_function="gelu", dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
        pad_token_id=50256, bos_token_id=50257, eos_token_id=50256, decoder_start_token_id=50258,
        suppress_tokens=None, begin_suppress_tokens=None)
    cfg._attn_implementation = "eager"
    m = WhisperForConditionalGeneration(cfg).eval()
    sd = {}
    t = lambda n: torch.from_numpy(T[n])

    def attn(src, dst):
        sd[dst + ".q_proj.weight"] = t(src + ".query.weight")
        sd[dst + ".q_proj.bias"] = t(src + ".query.bias")
        sd[dst + ".k_proj.weight"] = t(src + ".key.weight")
        sd[dst + ".v_proj.weight"] = t(src + ".value.weight")
        sd[dst + ".v_proj.bias"] = t(src + ".value.bias")
        sd[dst + ".out_proj.weight"] = t(src + ".out.weight")
        sd[dst + ".out_proj.bias"] = t(src + ".out.bias")

    def ln(src, dst):
        sd[dst + ".weight"] = t(src + ".weight")
        sd[dst + ".bias"] = t(src + ".bias")

    def mlp(src, dst):
        sd[dst + ".fc1.weight"] = t(src + ".mlp.0.weight")
        sd[dst + ".fc1.bias"] = t(src + ".mlp.0.bias")
        sd[dst + ".fc2.weight"] = t(src + ".mlp.2.weight")
        sd[dst + ".fc2.bias"] = t(src + ".mlp.2.bias")

    sd["model.encoder.conv1.weight"] = t("encoder.conv1.weight")
    sd["model.encoder.conv1.bias"] = t("encoder.conv1.bias").reshape(-1)
    sd["model.encoder.conv2.weight"] = t("encoder.conv2.weight")
    sd["model.encoder.conv2.bias"] = t("encoder.conv2.bias").reshape(-1)
    sd["model.encoder.embed_positions.weight"] = t("encoder.positional_embedding")
    for i in range(hp["n_audio_layer"]):
        s, d = "encoder.blocks.%d" % i, "model.encoder.layers.%d" % i
        ln(s + ".attn_ln", d + ".self_attn_layer_norm")
        attn(s + ".attn", d + ".self_attn")
        ln(s + ".mlp_ln", d + ".final_layer_norm")
        mlp(s, d)
    ln("encoder.ln_post", "model.encoder.layer_norm")
    sd["model.decoder.embed_tokens.weight"] = t("decoder.token_embedding.weight")
    sd["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
    sd["model.decoder.embed_positions.weight"] = t("decoder.positional_embedding")
    for i in range(hp["n_text_layer"]):
        s, d = "decoder.blocks.%d" % i, "model.decoder.layers.%d" % i
        ln(s + ".attn_ln", d + ".self_attn_layer_norm")
        attn(s + ".attn", d + ".self_attn")
        ln(s + ".cross_attn_ln", d + ".encoder_attn_layer_norm")
        attn(s + ".cross_attn", d + ".encoder_attn")
        ln(s + ".mlp_ln", d + ".final_layer_norm")
        mlp(s, d)
    ln("decoder.ln", "model.decoder.layer_norm")
    missing, unexpected = m.load_state_dict(sd, strict=False)
    missing = [k for k in missing if "k_proj.bias" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    return m, hp, filters


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    path = "/tmp/sw_golden_micro.bin"
    info = gen_model.generate(path, **MODEL_ARGS)
    torch.set_grad_enabled(False)
    m, hp, filters = hf_model(path)
    pcm16 = synth_audio.utterance(1, 0)
    pcm = synth_audio.to_f32(pcm16)

    # ---- log-mel: HF feature extractor with the file's own filterbank
    from transformers import WhisperFeatureExtractor
    fe = WhisperFeatureExtractor(feature_size=hp["n_mels"])
    fe.mel_filters = filters.T.astype(np.float64)
    hf_mel = fe(pcm, sampling_rate=16000, return_tensors="np")["input_features"][0]  # [80][3000]

    o = ora.Oracle(path, weight_round=False, act_round=ora.ACT_F32, gelu_erf=True)
    o_mel, n_org = o.mel(pcm)
    d_mel = np.abs(o_mel[:, :2998] - hf_mel[:, :2998])
    print("mel: max abs diff vs HF (frames 0..2997): %.3e  (values in [%.2f, %.2f])" %
          (d_mel.max(), hf_mel.min(), hf_mel.max()))
    assert d_mel.max() < 2e-3

    # ---- encoder (same mel window into both)
    win = o_mel[:, :3000]
    hf_enc = m.model.encoder(torch.from_numpy(win[None])).last_hidden_state[0].numpy()
    o_enc = o.encode(win)
    rel = np.abs(o_enc - hf_enc).max() / np.abs(hf_enc).max()
    print("encoder: max rel diff vs HF %.3e" % rel)
    assert rel < 2e-4

    # ---- decoder, teacher forced on the scripted transcript
    sp = info["special"]
    toks = np.array([sp["sot"], sp["sot"] + 1, sp["transcribe"]] + info["script"][:20], np.int32)
    hf_logits = m(encoder_outputs=(torch.from_numpy(hf_enc[None]),),
                  decoder_input_ids=torch.from_numpy(toks[None].astype(np.int64))).logits[0].numpy()
    o_logits = o.decode(toks, 0)
    rel_l = np.abs(o_logits - hf_logits).max() / np.abs(hf_logits).max()
    print("decoder logits: max rel diff vs HF %.3e; argmax agree %d/%d" %
          (rel_l, (o_logits.argmax(1) == hf_logits.argmax(1)).sum(), len(toks)))
    assert rel_l < 2e-4
    # incremental decoding == teacher forcing (KV cache correctness)
    inc = np.concatenate([o.decode(toks[:5], 0, slot=1)] +
                         [o.decode(toks[i:i + 1], i, slot=1) for i in range(5, len(toks))])
    assert np.abs(inc - o_logits).max() < 1e-3 * np.abs(o_logits).max()

    cols = np.r_[0:64, 50200:50464, 51800:51865]
    np.savez_compressed(
        os.path.join(out_dir, "micro_hf.npz"),
        model_args=np.array(repr(MODEL_ARGS)),
        pcm_config=np.array([1, 0]),
        hf_mel_sub=hf_mel[::8, ::10].astype(np.float32),
        hf_enc_sub=hf_enc[::15, :].astype(np.float32),
        tokens=toks,
        logit_cols=cols.astype(np.int32),
        hf_logits_sub=hf_logits[:, cols].astype(np.float32),
        hf_argmax=hf_logits.argmax(1).astype(np.int32))
    print("wrote", os.path.join(out_dir, "micro_hf.npz"))


if __name__ == "__main__":
    main()
