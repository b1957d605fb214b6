"""ncu target: a handful of skinny-GEMM launches (FC1 shape) with rotating weights."""
import ctypes, importlib.util, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("sw_binding", os.path.join(ROOT, "sentiric-stt-whisper-service_b200", "sw_binding.py"))
swb = importlib.util.module_from_spec(spec); spec.loader.exec_module(swb)
L = swb.lib()
vp, ci = ctypes.c_void_p, ctypes.c_int
L.sw_dev_skinny_gemm.argtypes = [vp, vp, ci, ci, ci, vp, ci, vp, vp, ci, vp]
R, N, K = 64, 5120, 1280
X = (torch.randn(R, K, device="cuda") * 0.5).bfloat16()
Ws = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(12)]
out = torch.empty(R, N, device="cuda", dtype=torch.bfloat16)
part = torch.zeros(8, R, 1280, device="cuda")
for i in range(12):
    L.sw_dev_skinny_gemm(X.data_ptr(), Ws[i].data_ptr(), R, N, K, None, 0, out.data_ptr(), None, 1, None)
W2 = [(torch.randn(1280, 5120, device="cuda") * 0.05).bfloat16() for _ in range(6)]
X2 = (torch.randn(R, 5120, device="cuda") * 0.5).bfloat16()
for i in range(6):
    L.sw_dev_skinny_gemm(X2.data_ptr(), W2[i].data_ptr(), R, 1280, 5120, None, 0, None, part.data_ptr(), 4, None)
torch.cuda.synchronize()
print("ok")
