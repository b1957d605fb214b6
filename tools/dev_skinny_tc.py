"""Development micro-benchmark (GPU box): the two decoder GEMM kernels (1 = mma.sync tiles, skinny_gemm.cu; 2 = tcgen05
with the weight rows in the M dimension, skinny_gemm_tc.cu) on the large-v3 decoder-step shapes, replayed from a CUDA
graph with rotating weights (they stream from HBM), alone and beside an occupier that holds 96 SMs (what a resident
cross attention of the other lane leaves: tools/dev_chain_occupied.py)."""
import ctypes, importlib.util, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("sw_binding", os.path.join(ROOT, "sentiric-stt-whisper-service_b200", "sw_binding.py"))
swb = importlib.util.module_from_spec(spec); spec.loader.exec_module(swb)
L = swb.lib()
vp, ci = ctypes.c_void_p, ctypes.c_int
L.sw_dev_skinny_gemm_k.argtypes = [ci, vp, vp, ci, ci, ci, vp, ci, vp, vp, ci, vp]
L.sw_dev_occupy.argtypes = [ci, ci, ctypes.c_float, ctypes.c_size_t]


def graph_time(fn, held, n_in_graph=64, reps=5):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(3): fn(i, st.cuda_stream)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for i in range(n_in_graph): fn(i, st.cuda_stream)
        g.replay(); st.synchronize()
        if held:
            assert L.sw_dev_occupy(held, 200 * 1024, 100.0, 0) == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): g.replay()
        e1.record(st); st.synchronize()
        if held:
            assert L.sw_dev_occupy(0, 0, 0.0, 0) == 0
    return e0.elapsed_time(e1) / (reps * n_in_graph) * 1e3


def run(R, N, K, split, n_w=16):
    X = (torch.randn(R, K, device="cuda") * 0.5).bfloat16()
    Ws = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(n_w)]
    bias = torch.randn(N, device="cuda")
    out = torch.empty(R, N, device="cuda", dtype=torch.bfloat16)
    res = dict(R=R, N=N, K=K)
    for kernel in (1, 2):
        sp = split if split > 0 else L.sw_dev_skinny_split_k(kernel, N, K)
        if split < 0:
            sp = 1
        part = torch.zeros(max(sp, 1), R, N, device="cuda")
        def fn(i, s):
            w = Ws[i % n_w]
            if sp == 1:
                rc = L.sw_dev_skinny_gemm_k(kernel, X.data_ptr(), w.data_ptr(), R, N, K, bias.data_ptr(), 0, out.data_ptr(), None, 1, s)
            else:
                rc = L.sw_dev_skinny_gemm_k(kernel, X.data_ptr(), w.data_ptr(), R, N, K, None, 0, None, part.data_ptr(), sp, s)
            assert rc == 0, swb.last_error()
        for held in (0, 96):
            res["k%d_split%d_held%d_us" % (kernel, sp, held)] = round(graph_time(fn, held), 2)
    print(json.dumps(res), flush=True)


d = 1280
for R in (64, 320):
    for (N, K, sp) in ((3 * d, d, -1), (d, d, 0), (4 * d, d, -1), (d, 4 * d, 0)):
        run(R, N, K, sp)
