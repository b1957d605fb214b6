import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "sentiric-stt-whisper-service_b200")
MODEL_DIR = os.environ.get("SW_TEST_MODEL_DIR", "/tmp/sw_test_models")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_pkg_module(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def swb():
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(PKG, "libsw_whisper.so")):
        ge.build()
    return load_pkg_module("sw_binding")


@pytest.fixture(scope="session")
def ora():
    from oracle import ora as o
    o.build()
    return o


def model_file(size, seed=1234, script_len=40, **kw):
    """Seeded synthetic ggml model, generated once per machine."""
    from tools import gen_model
    os.makedirs(MODEL_DIR, exist_ok=True)
    tag = "_".join("%s%s" % (k, v) for k, v in sorted(kw.items()))
    path = os.path.join(MODEL_DIR, "%s_s%d_n%d%s.bin" % (size, seed, script_len, ("_" + tag) if tag else ""))
    info_path = path + ".json"
    import json
    if not (os.path.exists(path) and os.path.exists(info_path)):
        info = gen_model.generate(path + ".tmp", size, seed=seed, script_len=script_len, **kw)
        os.replace(path + ".tmp", path)
        info["path"] = path
        json.dump(info, open(info_path, "w"))
    return path, json.load(open(info_path))


@pytest.fixture(scope="session")
def micro_model():
    return model_file("micro")


@pytest.fixture(scope="session")
def tiny_model():
    return model_file("tiny")


def seg_ids(result):
    return [t["id"] for s in result["segments"] for t in s["tokens"]]


def diff(a, b, path=""):
    """Where two nested results (dicts / lists / scalars) differ, as readable paths."""
    out = []
    if isinstance(a, dict):
        for k in a:
            out += diff(a[k], b[k], path + "/" + str(k))
    elif isinstance(a, list):
        if len(a) != len(b):
            return [path + " len %d vs %d" % (len(a), len(b))]
        for i, (x, y) in enumerate(zip(a, b)):
            out += diff(x, y, path + "[%d]" % i)
    elif a != b:
        out.append("%s: %r vs %r" % (path, a, b))
    return out
