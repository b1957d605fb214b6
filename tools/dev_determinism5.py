"""Development check (GPU box): the skinny GEMM and the tcgen05 GEMM in isolation, fixed inputs, repeated while a
second engine runs whole transcriptions from another host thread: bit-identical?"""
import ctypes, os, sys, threading
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import model_file
from tools import synth_audio
from tools.dev_determinism import swb
L = swb.lib()
vp, ci = ctypes.c_void_p, ctypes.c_int
L.sw_dev_skinny_gemm.argtypes = [vp, vp, ci, ci, ci, vp, ci, vp, vp, ci, vp]
L.sw_dev_gemm_bf16.argtypes = [vp] * 5 + [ci] * 8 + [vp]

path, info = model_file("small-4l", script_len=48)
clips = [synth_audio.utterance(3, i) for i in range(16)]
b = swb.Engine(path, max_batch=32, max_beams=5, n_lanes=1)
pb = b.default_params(0, language="en", temperature_inc=0.0, suppress_nst=1, token_timestamps=1)
stop = False
def hammer():
    while not stop:
        b.full_batch_pcm16(clips, pb)
st = torch.cuda.Stream()
torch.manual_seed(0)
cases = []
for (R, N, K, sp) in ((16, 3072, 768, 1), (64, 3072, 768, 1), (16, 5120, 1280, 1), (16, 1280, 5120, 0), (16, 3072, 1536, 2)):
    X = (torch.randn(R, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    split = sp if sp > 0 else L.sw_dev_skinny_split(N, K)
    cases.append((R, N, K, split, X, W, bias))
def run_case(c):
    R, N, K, split, X, W, bias = c
    with torch.cuda.stream(st):
        out = torch.zeros(R, N, device="cuda", dtype=torch.bfloat16)
        part = torch.zeros(max(split, 1), R, N, device="cuda")
        if split == 1:
            rc = L.sw_dev_skinny_gemm(X.data_ptr(), W.data_ptr(), R, N, K, bias.data_ptr(), 1, out.data_ptr(), None, 1, st.cuda_stream)
        else:
            rc = L.sw_dev_skinny_gemm(X.data_ptr(), W.data_ptr(), R, N, K, None, 0, None, part.data_ptr(), split, st.cuda_stream)
        assert rc == 0, swb.last_error()
        lg = torch.zeros(R, 4096, device="cuda")
        L.sw_dev_gemm_bf16(X.data_ptr(), W.data_ptr(), lg.data_ptr(), None, None, R, min(N, 4096), K, K, K, 4096, 2, 0, st.cuda_stream)
        st.synchronize()
    return (out.float() if split == 1 else part.sum(0)).cpu().numpy().copy(), part.cpu().numpy().copy(), lg.cpu().numpy().copy()
refs = [run_case(c) for c in cases]
th = threading.Thread(target=hammer); th.start()
bad = [[0, 0] for _ in cases]
for rep in range(400):
    for i, c in enumerate(cases):
        o, p, lg = run_case(c)
        mism = (p != refs[i][1]).any() or (o != refs[i][0]).any()
        if mism and bad[i][0] < 3:
            src, ref = (o, refs[i][0]) if c[3] == 1 else (p, refs[i][1])
            w = np.argwhere(src != ref)
            print("  case", c[:4], "rep", rep, "n_bad", len(w), "rows", sorted(set(w[:, -2].tolist()))[:20], "cols", sorted(set(w[:, -1].tolist()))[:48],
                  "max abs err", float(np.abs(src - ref).max()), "ref mag", float(np.abs(ref).mean()), flush=True)
        bad[i][0] += int(mism)
        bad[i][1] += int((lg != refs[i][2]).any())
stop = True
th.join()
for c, bd in zip(cases, bad):
    print("R,N,K,split", c[:4], "skinny mismatching reps", bd[0], "tcgen05 mismatching reps", bd[1], "of 400", flush=True)
b.close()
