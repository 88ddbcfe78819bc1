"""Fused out-proj / FFN2 + residual LayerNorm (care_gemm_add_ln) against the unfused care_gemm + care_add_ln pair,
CUDA events, buffers rotated so that inputs do not sit in L2.  Runs on the GPU box."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from care_b200 import _lib  # noqa: E402

F32, H16 = 0, 2


def timed(fn, iters=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    lib = _lib.load("fp16")
    h = ctypes.c_void_p()
    _lib.check(lib.care_ctx_create(ctypes.byref(h), 0), "ctx")
    st = torch.cuda.current_stream().cuda_stream
    for M, N, K in [(20480, 1024, 1024), (20480, 1024, 4096), (2560, 1024, 1024), (2560, 1024, 4096), (10240, 1024, 1024),
                    (20480, 512, 512), (20480, 768, 768),
                    (20480, 1024, 64), (20480, 1024, 256), (2560, 1024, 64)]:   # K = 64: the epilogue alone
        nbuf = 6 if M > 4096 else 24
        A = [torch.randn(M, K, device="cuda").half() for _ in range(nbuf)]
        res = [torch.randn(M, N, device="cuda").half() for _ in range(nbuf)]
        res32 = [torch.randn(M, N, device="cuda") for _ in range(nbuf)]
        W = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
        bias, gamma, beta = (torch.randn(N, device="cuda") for _ in range(3))
        y32 = torch.empty(M, N, device="cuda")
        out = [torch.empty(M, N, device="cuda", dtype=torch.float16) for _ in range(nbuf)]
        out32 = [torch.empty(M, N, device="cuda") for _ in range(nbuf)]

        def unfused(i):
            j = i % nbuf
            _lib.check(lib.care_gemm(h, H16, A[j].data_ptr(), K, W.data_ptr(), K, bias.data_ptr(), y32.data_ptr(), N, F32,
                                     M, N, K, 0, st), "gemm")
            _lib.check(lib.care_add_ln(h, H16, y32.data_ptr(), res[j].data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-12,
                                       M, N, out[j].data_ptr(), st), "add_ln")

        def fused(i):
            j = i % nbuf
            _lib.check(lib.care_gemm_add_ln(h, A[j].data_ptr(), K, W.data_ptr(), K, bias.data_ptr(), res[j].data_ptr(), H16,
                                            gamma.data_ptr(), beta.data_ptr(), 1e-12, out[j].data_ptr(), None, M, N, K, st),
                       "fused")

        def fused32(i):
            j = i % nbuf
            _lib.check(lib.care_gemm_add_ln(h, A[j].data_ptr(), K, W.data_ptr(), K, bias.data_ptr(), res32[j].data_ptr(),
                                            F32, gamma.data_ptr(), beta.data_ptr(), 1e-12, out[j].data_ptr(),
                                            out32[j].data_ptr(), M, N, K, st), "fused32")

        unfused(0)
        flops = 2.0 * M * N * K
        for name, fn, pair in (("gemm + add_ln", unfused, 0), ("fused (16-bit residual)", fused, 0),
                               ("fused, CTA pairs (16-bit residual)", fused, 1), ("fused (fp32 residual)", fused32, 0),
                               ("fused, CTA pairs (fp32 residual)", fused32, 1)):
            _lib.check(lib.care_ctx_set_option(h, b"gemm_ln_pair", pair), "option")
            ms = timed(fn)
            print("M=%5d N=%d K=%d  %-36s %.3f ms  %.0f TFLOP/s" % (M, N, K, name, ms, flops / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()
