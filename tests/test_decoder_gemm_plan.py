"""Host-side plans of the decoder's weight-streaming GEMMs (no GPU needed: the planners are plain functions of the
shape, exported through the development hooks of the C ABI). The tcgen05 kernel (skinny_gemm_tc.cu) serves d >= 768;
its K split must divide the k-blocks, leave >= 4 k-blocks per CTA and stay near the ~52 SMs a resident cross
attention of the other lane leaves; below d = 768 the mma.sync kernel's plan applies."""
import ctypes

import pytest

WIDTHS = {"tiny": 384, "base": 512, "small": 768, "medium": 1024, "large-v3": 1280}


@pytest.fixture
def lib(swb):
    L = swb.lib()
    L.sw_dev_skinny_split_k.argtypes = [ctypes.c_int] * 3
    L.sw_dev_skinny_split_k.restype = ctypes.c_int
    return L


@pytest.mark.parametrize("name", sorted(WIDTHS))
def test_tcgen05_plan_of_every_whisper_width(lib, name):
    d = WIDTHS[name]
    for N, K in ((3 * d, d), (d, d), (4 * d, d), (d, 4 * d)):
        s = lib.sw_dev_skinny_split_k(2, N, K)   # the tcgen05 kernel's plan whatever the engine would pick
        n_kb, tiles = K // 64, (N + 127) // 128
        assert 1 <= s <= 32 and n_kb % s == 0
        assert s == 1 or (n_kb // s >= 4 and tiles * s <= 56)
        # no larger admissible split was passed over
        assert not any(n_kb % t == 0 and n_kb // t >= 4 and tiles * t <= 56 for t in range(s + 1, 33))


def test_engine_routes_wide_models_to_the_tcgen05_kernel(lib):
    # kernel 0 = the engine's choice: it must equal the tcgen05 plan from d = 768 and the mma.sync plan below
    for name, d in WIDTHS.items():
        for N, K in ((d, d), (d, 4 * d)):
            want = lib.sw_dev_skinny_split_k(2 if d >= 768 else 1, N, K)
            assert lib.sw_dev_skinny_split_k(0, N, K) == want, (name, N, K)


def test_large_v3_plan_is_the_measured_one(lib):
    d = 1280
    assert lib.sw_dev_skinny_split_k(2, d, d) == 5        # 10 weight tiles x 5 slices = 50 CTAs of 4 k-blocks
    assert lib.sw_dev_skinny_split_k(2, d, 4 * d) == 5    # FC2: 50 CTAs of 16 k-blocks
    assert lib.sw_dev_skinny_split_k(2, 3 * d, d) == 1    # QKV: 30 tiles, bf16 out with bias
    assert lib.sw_dev_skinny_split_k(2, 4 * d, d) == 1    # FC1: 40 tiles, GELU in the epilogue
