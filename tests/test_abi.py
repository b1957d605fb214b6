"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol
include/sw_whisper.h declares, and fails loudly (no CPU fallback) without an sm_100 device."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sw_whisper.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"SW_API[^;(]*?\b(sw_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(swb):
    L = swb.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert set(syms) == set(swb.EXPORTS), set(syms) ^ set(swb.EXPORTS)


def test_every_entry_point_cites_the_reference(swb):
    src = open(os.path.join(ROOT, "include", "sw_whisper.h")).read()
    for needle in ("stt_engine.cpp:33", "stt_engine.cpp:245", "stt_engine.cpp:214", "stt_engine.cpp:261-292",
                   "main.cpp:71", "stt_engine.cpp:117"):
        assert needle in src, needle


def test_default_params_match_upstream_defaults(swb):
    L = swb.lib()
    g = L.sw_full_default_params(0)
    assert (g.strategy, g.best_of, g.beam_size) == (0, 5, -1)
    assert abs(g.temperature_inc - 0.2) < 1e-7 and abs(g.entropy_thold - 2.4) < 1e-6
    assert abs(g.logprob_thold + 1.0) < 1e-7 and abs(g.no_speech_thold - 0.6) < 1e-6
    assert g.suppress_blank == 1 and g.no_context == 1 and abs(g.max_initial_ts - 1.0) < 1e-7
    b = L.sw_full_default_params(1)
    assert (b.strategy, b.beam_size) == (1, 5)
    c = L.sw_ctx_default_params()
    assert (c.max_batch, c.max_beams) == (64, 5)
    assert L.sw_lang_id(b"en") == 0 and L.sw_lang_id(b"tr") == 9 and L.sw_lang_id(b"xx") == -1
    assert L.sw_version().startswith(b"sw_whisper")


def test_struct_layouts_match_the_header(swb):
    # sizes the C compiler gives the header's structs (guards the ctypes mirror against drift)
    import subprocess, tempfile
    code = r'''
    #include <stdio.h>
    #include "sw_whisper.h"
    int main(void){printf("%zu %zu %zu %zu %zu\n", sizeof(sw_ctx_params), sizeof(sw_model_info),
      sizeof(sw_full_params), sizeof(sw_token_data), sizeof(sw_stats)); return 0;}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(code)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"),
                               "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).split()
    want = [C.sizeof(swb.CtxParams), C.sizeof(swb.ModelInfo), C.sizeof(swb.FullParams),
            C.sizeof(swb.TokenData), C.sizeof(swb.Stats)]
    assert [int(x) for x in out] == want


def test_no_cpu_fallback(swb, micro_model):
    L = swb.lib()
    if L.sw_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError) as ei:
        swb.Engine(micro_model[0])
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_path_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "sentiric-stt-whisper-service_b200")
    for dp, _, fs in os.walk(pkg):
        if "build" in dp.split(os.sep):
            continue
        for f in fs:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "whisper_oracle" not in txt and "oracle/" not in txt and "from oracle" not in txt, f


def test_issuing_threads_are_elected_not_lane_tested(swb):
    """Regression guard for a performance bug found in SASS ('elect.sync'): a TMA producer or MMA issuer
    written as `if (lane == 0)` makes ptxas wrap every UTMALDG / UTCHMMA in a loop over the active lanes
    (R2UR ... ELECT ... BRA.U.ANY, ~110 cycles per instruction). With elect.sync the built objects contain the
    tensor-core / TMA instructions and none of those loops."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    build = os.path.join(ROOT, "sentiric-stt-whisper-service_b200", "build")
    want = {"gemm_tcgen05.o": ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG"), "attn_enc_tc.o": ("UTCHMMA", "UTMALDG"),
            "skinny_gemm.o": ("UTMALDG",), "xattn.o": ("UTMALDG",)}
    for obj, mnemonics in want.items():
        path = os.path.join(build, obj)
        if not os.path.exists(path):
            pytest.skip("objects not built in-tree (%s)" % obj)
        sass = subprocess.run([cuobjdump, "-sass", path], capture_output=True, text=True, timeout=300).stdout
        for m in mnemonics:
            assert m in sass, (obj, m)
        assert sass.count("BRA.U.ANY") == 0, "%s: %d lane loops around uniform-datapath instructions" % (
            obj, sass.count("BRA.U.ANY"))
