"""Quantised ggml files (a9: ModelManager / model load). whisper.cpp's quantize tool stores the 2-D weight
matrices as q4_0 / q4_1 / q5_0 / q5_1 / q8_0 blocks (header ftype 2 / 3 / 8 / 9 / 7 plus the quantisation
version in the thousands); the loader dequantises them at load. Checked against the independent `gguf`
package: a quantised file and its "twin" (the same model with gguf's DEQUANTISED values stored as f32) must
load to the very same bf16 weights, hence bit-identical logits and transcripts."""
import numpy as np
import pytest

from conftest import model_file, seg_ids
from tools import gen_model, ggml_io, synth_audio

pytest.importorskip("gguf")
QUANTS = ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0"]
GREEDY = dict(language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=1)


def test_generator_writes_block_quantised_matrices():
    """Header ftype, tensor types and block sizes as ggml defines them; the test-side reader (gguf's
    dequantisers) gets the twin's values back from the quantised file."""
    path, _ = model_file("micro", script_len=40, quant="q5_0")
    twin, _ = model_file("micro", script_len=40, quant="q5_0", quant_twin=True)
    hp, _, _, tq = ggml_io.read_ggml(path)
    hpt, _, _, tt = ggml_io.read_ggml(twin)
    assert hp["ftype"] == 2008 and hpt["ftype"] == 1
    w = "decoder.blocks.0.mlp.0.weight"
    assert np.array_equal(tq[w], tt[w])
    plain, _ = model_file("micro", script_len=40)
    _, _, _, tp = ggml_io.read_ggml(plain)
    err = np.abs(tq[w] - tp[w]).max()
    assert 0 < err < 0.02 * np.abs(tp[w]).max() * 4      # 5-bit codes: lossy, but close
    assert np.array_equal(tq["encoder.conv1.weight"], tp["encoder.conv1.weight"])   # 3-D: stays f16
    import os
    assert os.path.getsize(path) < 0.5 * os.path.getsize(plain)


@pytest.mark.gpu
@pytest.mark.parametrize("q", QUANTS)
def test_quantised_file_loads_to_the_same_weights_as_its_dequantised_twin(swb, ora, q):
    path, info = model_file("tiny", script_len=40, keyed=4, quant=q)
    twin, _ = model_file("tiny", script_len=40, keyed=4, quant=q, quant_twin=True)
    k = info["keyed"]
    clips = [synth_audio.keyed_clip(k, synth_audio.keyed_symbols(k, s), seed=s) for s in (300, 301)]
    sp = info["special"]
    toks = np.array([[sp["sot"], sp["sot"] + 1, sp["transcribe"]] + info["script"][:12]] * 2, np.int32)
    outs = []
    for p in (path, twin):
        e = swb.Engine(p, max_batch=4)
        assert e.info.ftype == (2000 + gen_model.QUANT_FTYPE[q] if p == path else 1)
        res = e.full_batch_pcm16(clips, e.default_params(0, **GREEDY))
        logits = e.decode_logits(toks)   # over the cross-KV the run above left behind
        outs.append((res, logits))
        e.close()
    assert outs[0][0] == outs[1][0]
    assert np.array_equal(outs[0][1], outs[1][1])
    # the quantised keyed model still listens, and the CUDA path agrees with the oracle run on the twin
    from test_gpu_parity import compare_results
    o = ora.Oracle(twin, weight_round=False, act_round=ora.ACT_F16)
    for c, s, r in zip(clips, (300, 301), outs[0][0]):
        assert seg_ids(r) == gen_model.keyed_expected_tokens(info, synth_audio.keyed_symbols(k, s))
        compare_results(r, o.full(synth_audio.to_f32(c), o.default_params(0, **GREEDY)))


@pytest.mark.gpu
def test_k_quants_are_rejected_with_a_clear_error(swb, tmp_path):
    path, _ = model_file("micro", script_len=40, quant="q8_0")
    data = bytearray(open(path, "rb").read())
    bad = tmp_path / "kquant.bin"
    data[4 + 40: 4 + 44] = (2012).to_bytes(4, "little")   # ftype: mostly q3_k
    bad.write_bytes(bytes(data))
    with pytest.raises(RuntimeError) as ei:
        swb.Engine(str(bad))
    assert "ftype" in str(ei.value) and "not supported" in str(ei.value)
