"""Development check (GPU box): is a beam-search batch bit-identical across reruns, with 1 and 2 lanes?"""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import model_file
from tools import synth_audio
spec = importlib.util.spec_from_file_location("sw_binding", os.path.join(ROOT, "sentiric-stt-whisper-service_b200", "sw_binding.py"))
swb = importlib.util.module_from_spec(spec); spec.loader.exec_module(swb)

def diff(a, b, path=""):
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

def main():
    path, info = model_file("small-4l", script_len=48)
    clips = [synth_audio.utterance(3, i) for i in range(32)]
    for strategy, kw in ((1, dict(language="en", temperature_inc=0.0, suppress_nst=1, beam_size=5)),
                         (0, dict(language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=1))):
        for lanes in (1, 2):
            cases = os.environ.get("CASES")  # e.g. "1:1,0:2" = (strategy:lanes) pairs to run
            if cases and "%d:%d" % (strategy, lanes) not in cases.split(","):
                continue
            e = swb.Engine(path, max_batch=32, max_beams=5, n_lanes=lanes)
            pe = e.default_params(strategy, **kw)
            ref = e.full_batch_pcm16(clips, pe)
            for rep in range(4):
                got = e.full_batch_pcm16(clips, pe)
                d = diff(ref, got)
                print("strategy", strategy, "lanes", lanes, "rep", rep, "diffs", len(d), d[:4], flush=True)
            e.close()


if __name__ == "__main__":
    main()
