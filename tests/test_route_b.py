"""INTEGRATION.md route B, built for real: the reference's OWN src/stt_engine.cpp (+ prosody_extractor.cpp,
speaker_cluster.cpp), unmodified and compiled where it lies (oracle/Makefile target `routeb` ->
oracle/_ref/route_b_cli), against include/compat/whisper.h + samplerate.h and linked to
libwhisper_compat.so -> libsw_whisper.so. CPU part: the shim exports what the reference's translation unit
imports, and the reference's own constructor error surfaces without a GPU. GPU part: that binary and the
route-A facade (host/stt_engine.cpp through stt_cli) return the same TranscriptionResults for the same clip."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG, ROOT, model_file
from tools import synth_audio

CLI_B = os.path.join(ROOT, "oracle", "_ref", "route_b_cli")
COMPAT = os.path.join(PKG, "libwhisper_compat.so")

# the whisper.cpp symbols stt_engine.cpp binds (+ whisper_log_set from main.cpp:71) and
# libsamplerate's src_simple (stt_engine.cpp:103)
REFERENCE_IMPORTS = """whisper_context_default_params whisper_init_from_file_with_params whisper_init_state
whisper_free_state whisper_free whisper_vad_default_context_params whisper_vad_init_from_file_with_params
whisper_vad_detect_speech whisper_vad_free whisper_full_default_params whisper_full_with_state
whisper_full_n_segments_from_state whisper_full_get_segment_text_from_state whisper_full_get_segment_t0_from_state
whisper_full_get_segment_t1_from_state whisper_full_get_segment_speaker_turn_next_from_state
whisper_full_n_tokens_from_state whisper_full_get_token_data_from_state whisper_token_to_str whisper_token_eot
whisper_log_set src_simple""".split()


def build():
    subprocess.check_call(["make", "-s", "-C", PKG])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "routeb"])


def dyn_symbols(path, undefined):
    out = subprocess.run(["nm", "-D", "--undefined-only" if undefined else "--defined-only", path],
                         capture_output=True, text=True, check=True).stdout
    return {l.split()[-1] for l in out.splitlines() if l.strip()}


def test_compat_library_exports_every_symbol_the_reference_binds():
    build()
    exported = dyn_symbols(COMPAT, undefined=False)
    missing = [s for s in REFERENCE_IMPORTS if s not in exported]
    assert not missing, missing
    hdr = open(os.path.join(ROOT, "include", "compat", "whisper.h")).read()
    for s in REFERENCE_IMPORTS[:-1]:
        assert s + "(" in hdr, s


def test_reference_translation_unit_links_against_the_shim():
    build()
    if not os.path.exists(CLI_B):
        pytest.skip("oracle/_ref/route_b_cli not built (no /root/reference or third-party headers here)")
    wanted = {s for s in dyn_symbols(CLI_B, undefined=True) if s.startswith("whisper_") or s.startswith("src_")}
    assert wanted and wanted <= set(REFERENCE_IMPORTS)          # nothing outside the documented surface
    assert wanted <= dyn_symbols(COMPAT, undefined=False)
    assert {"whisper_full_with_state", "whisper_init_from_file_with_params", "src_simple"} <= wanted
    ldd = subprocess.run(["ldd", CLI_B], capture_output=True, text=True).stdout
    assert "libwhisper_compat.so" in ldd and "libsw_whisper.so" in ldd and "not found" not in ldd


def test_reference_constructor_error_without_a_model(tmp_path):
    """stt_engine.cpp:34: a context that cannot be created makes the reference's own constructor throw."""
    build()
    if not os.path.exists(CLI_B):
        pytest.skip("oracle/_ref/route_b_cli not built")
    raw = tmp_path / "x.raw"
    np.zeros(16000, np.int16).tofile(raw)
    r = subprocess.run([CLI_B, str(tmp_path), "missing.bin", str(raw)], capture_output=True, text=True)
    assert r.returncode == 1 and "Whisper model initialization failed" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("beam,rate", [(1, 16000), (5, 16000), (1, 48000)])
def test_route_b_equals_route_a(tmp_path, beam, rate):
    from test_host_facade import HOST, build_host
    build_host()
    if not os.path.exists(CLI_B):
        pytest.skip("oracle/_ref/route_b_cli not built")
    path, info = model_file("tiny", script_len=40, keyed=4)
    k = info["keyed"]
    sy = synth_audio.keyed_symbols(k, 41)
    clip = synth_audio.keyed_clip(k, sy, seed=41, sr=rate)
    if rate != 16000:
        clip = np.round(np.clip(clip, -1, 1) * 32767.0).astype(np.int16)
    raw = tmp_path / "clip.raw"
    clip.tofile(raw)
    a = subprocess.run([os.path.join(HOST, "build", "stt_cli"), os.path.dirname(path), os.path.basename(path), str(raw),
                        "1", str(beam), "batch", str(rate)], capture_output=True, text=True, timeout=300)
    b = subprocess.run([CLI_B, os.path.dirname(path), os.path.basename(path), str(raw), str(beam), str(rate), "en"],
                       capture_output=True, text=True, timeout=300)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    ra = json.loads(a.stdout.strip().splitlines()[0])
    rb = json.loads(b.stdout.strip().splitlines()[-1])
    assert ra["token_count"] == rb["token_count"] > 20
    assert len(ra["segments"]) == len(rb["segments"]) >= 2
    for sa, sb in zip(ra["segments"], rb["segments"]):
        for key in ("t0", "t1", "text", "language", "speaker", "gender", "emotion"):
            assert sa[key] == sb[key], key
        assert [t[0] for t in sa["tokens"]] == [t[0] for t in sb["tokens"]]
        assert [t[2:] for t in sa["tokens"]] == [t[2:] for t in sb["tokens"]]          # token t0 / t1
        assert np.allclose([t[1] for t in sa["tokens"]], [t[1] for t in sb["tokens"]], atol=1e-2)
        assert abs(sa["prob"] - sb["prob"]) < 1e-2
        # route A computes prosody with the CUDA kernels, route B with the reference's own host code: bit-identical
        assert sa["prosody"] == sb["prosody"]
