"""Timeline of vocab_beam_2sm_kernel at a strong-scaling shard size (debug build with -DCARE_VB_TRACE):
`python scripts/vb_trace.py build` here (cross-compiles lib/libcare_b200_trace.so), `python scripts/vb_trace.py` on the
GPU box.  Prints, for a few CTA pairs, the SM-clock stamps of the leader CTA's roles relative to kernel entry."""
import ctypes
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIBDIR = os.path.join(ROOT, "care_b200", "lib")
TRACE_LIB = os.path.join(LIBDIR, "libcare_b200_trace.so")


def build():
    from care_b200 import build as b
    b.build()
    odir = os.path.join(LIBDIR, "obj_trace")
    os.makedirs(odir, exist_ok=True)
    obj = os.path.join(odir, "vocab_beam.o")
    subprocess.check_call([b.NVCC] + [f for f in b.FLAGS if f not in ("-Xptxas", "-v")] + ["-DCARE_VB_TRACE", "-c",
                          os.path.join(b.SRC, "vocab_beam.cu"), "-o", obj])
    others = [o for o in glob.glob(os.path.join(LIBDIR, "obj_fp16", "*.o")) if not o.endswith("vocab_beam.o")]
    subprocess.check_call([b.NVCC, "-shared", "-o", TRACE_LIB, obj] + others + ["-gencode", "arch=compute_100a,code=sm_100a"])
    print(TRACE_LIB)


def main():
    import numpy as np
    import torch
    from care_b200 import _lib
    _lib.LIB_PATH = TRACE_LIB
    lib = _lib.load("fp16")
    lib.care_debug_vb_trace.restype = ctypes.c_int
    lib.care_debug_vb_trace.argtypes = [ctypes.c_void_p]
    h = ctypes.c_void_p()
    _lib.check(lib.care_ctx_create(ctypes.byref(h), 0), "ctx")
    st = torch.cuda.current_stream().cuda_stream
    R, V, d, K = int(os.environ.get("VBT_ROWS", 2560)), 14745, 1024, 5
    x = torch.randn(R, d, device="cuda").half()
    W = (torch.randn(V, d, device="cuda") / d ** 0.5).half()
    nseg = lib.care_vocab_beam_nseg(h, R, V)
    partials = torch.empty(R * nseg * 14, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    buf = np.zeros((128, 5, 64), dtype=np.int64)
    for cold in (1, 0, 1):
        for rep in range(3):
            if cold:
                flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.care_vocab_beam_partials(h, x.data_ptr(), d, W.data_ptr(), d, R, V, d, K, partials.data_ptr(), nseg,
                                                    st), "vocab")
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        assert lib.care_debug_vb_trace(buf.ctypes.data) == 0
        print("==== %s L2, last of 3 launches: %.1f us (CUDA events)" % ("cold" if cold else "warm", ms * 1e3))
        for c in (0, 1, 36, 73):
            t = buf[c]
            t0 = t[0][0]
            rel = lambda a: [int(v - t0) if v else None for v in a]
            print("cluster %d: roles start +%d clk, exit +%d clk" % (c, t[0][1] - t0, t[0][2] - t0))
            print("  producer, first k-block of tile issued:", rel(t[1][:9]))
            m = rel(t[2][:27])
            print("  mma per tile (acc free, operands landed, last mma issued):", [tuple(m[3 * i:3 * i + 3]) for i in range(9)])
            for g in (0, 1):
                e = rel(t[3 + g][:10])
                print("  epilogue group %d per tile (acc full, fold done):" % g, [tuple(e[2 * i:2 * i + 2]) for i in range(5)])
        ends = [int(buf[c][0][2] - buf[c][0][0]) for c in range(74)]
        starts = [int(buf[c][0][1] - buf[c][0][0]) for c in range(74)]
        print("  all clusters: roles start min/max %d/%d clk, exit min/max %d/%d clk" % (min(starts), max(starts), min(ends), max(ends)))


if __name__ == "__main__":
    build() if sys.argv[1:] == ["build"] else main()
