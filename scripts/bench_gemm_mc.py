"""The layer GEMM shapes of cfg4 on the four care_gemm variants (single-CTA tiles, CTA pairs, clusters of 4 / 2 pairs with
the A tile multicast), CUDA events, rotating operand buffers.  Runs on the GPU box."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from care_b200 import _lib  # noqa: E402

F32, H16 = 0, 2


def timed(fn, iters=30, warm=4):
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
    lib.care_ctx_set_option(h, b"debug", 1)
    st = torch.cuda.current_stream().cuda_stream
    for M, N, K in [(20480, 3072, 1024), (20480, 4096, 1024), (20480, 1024, 1024), (8192, 8192, 8192), (466944, 2048, 1024),
                    (2560, 3072, 1024), (2560, 4096, 1024)]:
        nbuf = 2 if M > 100000 else (4 if M > 4096 else 16)
        A = [torch.randn(M, K, device="cuda").half() for _ in range(nbuf)]
        W = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
        bias = torch.randn(N, device="cuda")
        out = [torch.empty(M, N, device="cuda", dtype=torch.float16) for _ in range(nbuf)]
        flops = 2.0 * M * N * K
        for name, mode in (("single-CTA tiles", 0), ("CTA pairs", 1), ("4-pair clusters, A multicast", 4),
                           ("2-pair clusters, A multicast", 5)):
            _lib.check(lib.care_ctx_set_option(h, b"gemm_2sm", mode), "option")

            def run(i):
                j = i % nbuf
                _lib.check(lib.care_gemm(h, H16, A[j].data_ptr(), K, W.data_ptr(), K, bias.data_ptr(), out[j].data_ptr(), N,
                                         H16, M, N, K, 0, st), "gemm")

            ms = timed(run)
            print("M=%6d N=%5d K=%5d  %-30s %-34s %.3f ms  %.0f TFLOP/s" % (
                M, N, K, name, lib.care_ctx_last_kernel(h, b"gemm").decode(), ms, flops / ms / 1e9), flush=True)
        del A, out


if __name__ == "__main__":
    main()
