"""Cross-attention step kernel alone at the cfg4 shape (4096 videos x 16 heads, Lm = 114, beam 5), CUDA events, K/V buffers
rotated so that nothing is served from L2.  Runs on the GPU box."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from care_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load("fp16")
    h = ctypes.c_void_p()
    _lib.check(lib.care_ctx_create(ctypes.byref(h), 0), "ctx")
    st = torch.cuda.current_stream().cuda_stream
    B, K, H, d, Lm = 4096, 5, 16, 1024, 114
    R = B * K
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6450.0
    q = torch.randn(R, d, device="cuda").half()
    kvs = [torch.randn(B, Lm, 2 * d, device="cuda").half() for _ in range(3)]
    bias = torch.randn(H, Lm, device="cuda")
    done = torch.zeros(B, device="cuda", dtype=torch.int32)
    out = torch.zeros(R, d, device="cuda", dtype=torch.float16)

    def run(i):
        _lib.check(lib.care_cross_attn_step(h, 2, q.data_ptr(), d, kvs[i % 3].data_ptr(), Lm, B, K, H, d, bias.data_ptr(),
                                            done.data_ptr(), out.data_ptr(), st), "cross")

    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 30
    for i in range(n):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    nbytes = B * Lm * 2 * d * 2 + 2 * R * d * 2
    print("cross attention: %.4f ms per launch, %.0f GB/s = %.3f of the measured copy bandwidth %.0f" % (
        ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak, peak))


if __name__ == "__main__":
    main()
