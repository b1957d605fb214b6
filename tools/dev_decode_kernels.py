"""Development micro-benchmark (GPU box): decoder-step kernels in isolation, replayed from a CUDA
graph (no CPU launch overhead), weights rotated so they stream from HBM."""
import ctypes, importlib.util, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("sw_binding", os.path.join(ROOT, "sentiric-stt-whisper-service_b200", "sw_binding.py"))
swb = importlib.util.module_from_spec(spec); spec.loader.exec_module(swb)
L = swb.lib()
vp, ci = ctypes.c_void_p, ctypes.c_int
L.sw_dev_skinny_gemm.argtypes = [vp, vp, ci, ci, ci, vp, ci, vp, vp, ci, vp]
L.sw_dev_layer_norm.argtypes = [vp, ci, ci, vp, vp, vp, vp, ci, vp, vp]
L.sw_dev_gemm_bf16.argtypes = [vp] * 5 + [ci] * 8 + [vp]

def graph_time(fn, n_in_graph=64, reps=5):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(3): fn(i, st.cuda_stream)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for i in range(n_in_graph): fn(i, st.cuda_stream)
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): g.replay()
        e1.record(st); st.synchronize()
    return e0.elapsed_time(e1) / (reps * n_in_graph) * 1e3

def run(R, N, K, split, n_w=int(os.environ.get("N_W", "16"))):
    X = (torch.randn(R, K, device="cuda") * 0.5).bfloat16()
    Ws = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(n_w)]
    bias = torch.randn(N, device="cuda")
    out = torch.empty(R, N, device="cuda", dtype=torch.bfloat16)
    sp = split if split > 0 else L.sw_dev_skinny_split(N, K)
    part = torch.zeros(max(sp, 1), R, N, device="cuda")
    def fn(i, s):
        w = Ws[i % n_w]
        if sp == 1:
            rc = L.sw_dev_skinny_gemm(X.data_ptr(), w.data_ptr(), R, N, K, bias.data_ptr(), 0, out.data_ptr(), None, 1, s)
        else:
            rc = L.sw_dev_skinny_gemm(X.data_ptr(), w.data_ptr(), R, N, K, None, 0, None, part.data_ptr(), sp, s)
        assert rc == 0, swb.last_error()
    us = graph_time(fn)
    def fn2(i, s):
        w = Ws[i % n_w]
        L.sw_dev_gemm_bf16(X.data_ptr(), w.data_ptr(), out.data_ptr(), bias.data_ptr(), None, R, N, K, K, K, N, 0, 0, s)
    us2 = graph_time(fn2)
    print(json.dumps(dict(R=R, N=N, K=K, split=sp, us=round(us, 2), gbs=round(N * K * 2 / us / 1e3, 1), tcgen05_us=round(us2, 2))), flush=True)

d = 1280
if os.environ.get("SHAPES"):  # "N,K,split;N,K,split;..." at R = 64
    for t in os.environ["SHAPES"].split(";"):
        N, K, sp = (int(v) for v in t.split(","))
        run(64, N, K, sp)
    sys.exit(0)
for R in (64, 8):
    for (N, K, sp) in ((3 * d, d, 1), (3 * d, d, 2), (3 * d, d, 4), (d, d, 1), (d, d, 2), (d, d, 4), (4 * d, d, 1), (4 * d, d, 2), (d, 4 * d, 4), (d, 4 * d, 8), (d, 4 * d, 16)):
        run(R, N, K, sp)
x = torch.randn(64, d, device="cuda"); g = torch.ones(d, device="cuda"); b = torch.zeros(d, device="cuda")
out = torch.empty(64, d, device="cuda", dtype=torch.bfloat16); part = torch.randn(8, 64, d, device="cuda") * 0.01
print("LN plain us", graph_time(lambda i, s: L.sw_dev_layer_norm(x.data_ptr(), 64, d, g.data_ptr(), b.data_ptr(), out.data_ptr(), None, 0, None, s)))
print("LN fused(4) us", graph_time(lambda i, s: L.sw_dev_layer_norm(x.data_ptr(), 64, d, g.data_ptr(), b.data_ptr(), out.data_ptr(), part.data_ptr(), 4, b.data_ptr(), s)))
print("LN fused(8) us", graph_time(lambda i, s: L.sw_dev_layer_norm(x.data_ptr(), 64, d, g.data_ptr(), b.data_ptr(), out.data_ptr(), part.data_ptr(), 8, b.data_ptr(), s)))
