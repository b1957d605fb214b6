"""bench.py's output contract, checked on the CPU arm with a tiny model (seconds): one JSON line with the keys
the driver reads; the reference arm adds impl / cpu_baseline / a zero-copy e2e block."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "micro",
                        "--script-len", "24", "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["metric"].startswith("audio-sec/sec") and d["unit"] == "audio-sec/sec" and d["higher_is_better"] is True
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1


def test_bench_names_the_baseline_metric_and_never_imports_the_oracle_on_the_gpu_path():
    src = open(os.path.join(ROOT, "bench.py")).read()
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "audio-sec/sec (RTFx)" in src and "RTFx" in base["metric"]
    # the oracle is only imported inside the reference arm and the cpu_baseline legs
    for i, line in enumerate(src.splitlines()):
        if "from oracle import" in line:
            ctx = "\n".join(src.splitlines()[max(0, i - 12):i])
            assert "def run_reference" in src[:src.index(line)] and (
                "no_cpu_baseline" in ctx or "run_reference" in ctx or "cpu_baseline" in ctx or "impl = pro" in ctx), line


def test_our_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback on the measured path: without a CUDA device bench.py's own arm exits with an error and
    prints no JSON line (the reference arm is the only thing that runs on host cores)."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--model", "micro", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
