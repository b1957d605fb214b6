"""Development check for the tcgen05 GEMM (run on the GPU box).

Compares sw_dev_gemm_bf16 against a torch fp32 matmul of the same bf16 inputs and
times it with CUDA events. Not part of the product path.
"""
import ctypes, os, sys, time, json
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "sentiric-stt-whisper-service_b200", "libsw_whisper.so"))
lib.sw_last_error.restype = ctypes.c_char_p
lib.sw_dev_gemm_bf16.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 8 + [ctypes.c_void_p]
lib.sw_dev_gemm_bf16.restype = ctypes.c_int


def run(M, N, K, flags=0, bias=False, res=False, block_n=0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.5).bfloat16()
    out_f32 = bool(flags & 2)
    ldc = (N + 7) // 8 * 8
    Cfull = torch.full((M, ldc), 7.0, device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    C = Cfull[:, :N]
    bias_t = torch.randn(M if flags & 4 else N, device="cuda", generator=g) if bias else None
    res_full = torch.randn(M, ldc, device="cuda", generator=g) if res else None
    res_t = res_full[:, :N] if res else None
    rc = lib.sw_dev_gemm_bf16(A.data_ptr(), B.data_ptr(), Cfull.data_ptr(),
                              bias_t.data_ptr() if bias else None,
                              res_full.data_ptr() if res else None,
                              M, N, K, K, K, ldc, flags, block_n, None)
    if rc != 0:
        return dict(M=M, N=N, K=K, flags=flags, ok=False, err=lib.sw_last_error().decode())
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    if bias:
        ref = ref + (bias_t[:, None] if flags & 4 else bias_t[None, :])
    if flags & 1:
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    if res:
        ref = ref + res_t
    err = (C.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = 1e-3 * max(scale, 1.0) if out_f32 else 1e-2 * max(scale, 1.0)
    return dict(M=M, N=N, K=K, flags=flags, bias=bias, res=res, block_n=block_n,
                max_err=err, ref_max=scale, ok=bool(err <= tol))


def bench(M, N, K, block_n=0, iters=20):
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        lib.sw_dev_gemm_bf16(A.data_ptr(), B.data_ptr(), C.data_ptr(), None, None, M, N, K, K, K, N, 0, block_n, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.sw_dev_gemm_bf16(A.data_ptr(), B.data_ptr(), C.data_ptr(), None, None, M, N, K, K, K, N, 0, block_n, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # torch (cuBLAS) for context
    for _ in range(3):
        torch.matmul(A, B.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(A, B.t())
    e1.record()
    torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / iters
    fl = 2.0 * M * N * K
    return dict(M=M, N=N, K=K, block_n=block_n, ms=ms, tflops=fl / ms / 1e9, cublas_ms=ms_t,
                cublas_tflops=fl / ms_t / 1e9)


def bench_epilogue(M, N, K, flags, bias, res, iters=20, block_n=256):
    """The shapes as the encoder launches them: bias / GELU / f32 residual in and f32 out in the epilogue."""
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    out_f32 = bool(flags & 2)
    C = torch.zeros(M, N, device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    bias_t = torch.randn(N, device="cuda") if bias else None
    args = (A.data_ptr(), B.data_ptr(), C.data_ptr(), bias_t.data_ptr() if bias else None,
            C.data_ptr() if res else None, M, N, K, K, K, N, flags, block_n, None)  # residual in place, as the engine does
    for _ in range(3):
        lib.sw_dev_gemm_bf16(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.sw_dev_gemm_bf16(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return dict(M=M, N=N, K=K, flags=flags, bias=bias, res=res, block_n=block_n, ms=ms, tflops=2.0 * M * N * K / ms / 1e9)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    results = []
    cases = [
        (128, 64, 64, 0, False, False, 64),
        (128, 128, 64, 2, False, False, 128),
        (128, 256, 128, 2, False, False, 256),
        (256, 256, 256, 2, False, False, 0),
        (1500, 384, 384, 0, True, False, 0),
        (1500, 1536, 384, 1, True, False, 0),
        (1500, 384, 1536, 2, True, True, 0),
        (3000, 1280, 384, 1, True, False, 0),
        (777, 200, 136, 2, True, True, 0),      # ragged M/N/K tails
        (5000, 51866, 384, 2, False, False, 256),
        (96000, 1280, 1280, 0, True, False, 256),
        (4096, 4096, 4096, 2, False, False, 256),
        (1000, 640, 512, 6, True, False, 0),    # row bias
        (1000, 640, 512, 7, True, True, 0),     # row bias + GELU + residual, f32 out
        (3000, 1280, 1280, 2, True, True, 256),  # out-projection as the encoder launches it
        (3000, 5120, 1280, 1, True, False, 256),  # FC1: GELU, bf16 out
        (1501, 392, 200, 0, True, True, 0),     # ragged everything, bf16 out with residual
        # block_n = 512 forces the CTA-pair kernel (256 x 256 tile per two SMs)
        (256, 256, 64, 2, False, False, 512),
        (512, 512, 256, 2, False, False, 512),
        (3000, 1280, 1280, 2, True, True, 512),
        (3000, 5120, 1280, 1, True, False, 512),
        (777, 200, 136, 2, True, True, 512),
        (1501, 392, 200, 0, True, True, 512),
        (1000, 640, 512, 7, True, True, 512),
        (96000, 1280, 1280, 0, True, False, 512),
        (5000, 51866, 384, 2, False, False, 512),
    ]
    allok = True
    for c in cases:
        r = run(*c)
        print(json.dumps(r), flush=True)
        allok &= r["ok"]
    print("ALL_OK" if allok else "SOME_FAILED", flush=True)
    if allok:
        for (M, N, K, bn) in [(8192, 8192, 8192, 256), (96000, 1280, 1280, 256), (96000, 5120, 1280, 256),
                              (96000, 1280, 5120, 256), (96000, 3840, 1280, 256), (96000, 1280, 1280, 128),
                              (1500 * 32, 512, 512, 128), (1500 * 32, 512, 512, 256), (1500*32, 2048, 512, 256),
                              (8192, 8192, 8192, 512), (96000, 1280, 1280, 512), (96000, 5120, 1280, 512),
                              (96000, 1280, 5120, 512), (96000, 3840, 1280, 512)]:
            print(json.dumps(bench(M, N, K, bn)), flush=True)
        # out-projection (f32 residual stream in place), FC1 (GELU, bf16 out), FC2 (residual), QKV (bias)
        for rep in range(2):
            for (M, N, K, fl, bi, re) in [(96000, 1280, 1280, 2, True, True), (96000, 5120, 1280, 1, True, False),
                                          (96000, 1280, 5120, 2, True, True), (96000, 3840, 1280, 0, True, False),
                                          (96000, 2560, 1280, 0, True, False)]:
                for bn in (256, 512):  # 512 = CTA pairs
                    print(json.dumps(bench_epilogue(M, N, K, fl, bi, re, block_n=bn)), flush=True)
    sys.exit(0 if allok else 1)
