"""Development timing (GPU box): device time of one decoder step at large-v3 widths, 64 rows,
replayed from the step graph, through the teacher-forced sw_decode_logits hook (fixed number of
steps whatever the logits are). Runs 8-layer and 2-layer models so that the per-layer cost
((t8 - t2) / 6) separates from the per-step fixed part (embedding, final LN, logits GEMM).
`--sweep` repeats the measurement with SW_SKIP=<bit> for every kernel of the layer: the difference to
the baseline is that kernel's in-graph cost. Engine switches (SW_PDL=0, SW_GRAPHS=0) apply."""
import importlib.util, json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SKIPS = [(0, "baseline"), (1, "layer_norm x3"), (2, "qkv gemm"), (4, "self attention"), (8, "wo gemm"),
         (16, "wxq gemm"), (32, "reduce q"), (64, "cross attention + combine"), (256, "wxo gemm"),
         (512, "fc1 gemm"), (1024, "fc2 gemm"), (2047 - 64, "everything but cross attention")]


def run(size, n=64, n_tok=40):
    from tools import gen_model
    spec = importlib.util.spec_from_file_location("sw_binding", os.path.join(ROOT, "sentiric-stt-whisper-service_b200", "sw_binding.py"))
    swb = importlib.util.module_from_spec(spec); spec.loader.exec_module(swb)
    path = "/tmp/sw_%s_s0.bin" % size
    if not os.path.exists(path):
        gen_model.generate(path, size, seed=7)
    eng = swb.Engine(path, max_batch=n, max_beams=1)
    tok = np.random.default_rng(0).integers(0, 50000, size=(n, n_tok)).astype(np.int32)
    eng.decode_logits(tok)
    eng.stats(reset=True)
    eng.decode_logits(tok)
    st = eng.stats(reset=True)
    eng.close()
    return st["ms_decode"] / max(1, st["n_steps"])


def one():
    t8, t2 = run("large-v3-8l"), run("large-v3-2l")
    per_layer = (t8 - t2) / 6.0
    return dict(skip=int(os.environ.get("SW_SKIP", "0")), pdl=os.environ.get("SW_PDL", "1"),
                us_per_layer=round(per_layer * 1e3, 2), us_fixed=round((t2 - 2 * per_layer) * 1e3, 2),
                est_ms_step_32l=round(t2 + 30 * per_layer, 3))


if __name__ == "__main__":
    if "--sweep" in sys.argv:
        base = None
        for bit, name in SKIPS:
            env = dict(os.environ, SW_SKIP=str(bit))
            out = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
            try:
                r = json.loads(out.stdout.strip().splitlines()[-1])
            except Exception:
                print("skip", bit, name, "FAILED", out.stdout[-300:], out.stderr[-300:], flush=True)
                continue
            if bit == 0:
                base = r["us_per_layer"]
            r["kernel"] = name
            r["in_graph_cost_us"] = None if base is None else round(base - r["us_per_layer"], 2)
            print(json.dumps(r), flush=True)
    else:
        print(json.dumps(one()), flush=True)
