"""Builds the C-ABI libraries in-tree with nvcc for sm_100a (cross-compiles without a GPU):
care_b200/lib/libcare_b200.so (16-bit operand type = IEEE fp16, the default throughput mode) and
care_b200/lib/libcare_b200_bf16.so (same sources with -DCARE_USE_BF16)."""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libcare_b200.so")
LIB_BF16 = os.path.join(LIB_DIR, "libcare_b200_bf16.so")
VARIANTS = (("fp16", LIB, []), ("bf16", LIB_BF16, ["-DCARE_USE_BF16"]))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _sources():
    return sorted(glob.glob(os.path.join(SRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    for p in sorted(glob.glob(os.path.join(SRC, "*")) + [os.path.join(os.path.dirname(HERE), "include", "care_b200.h")]):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "build.sha256")
    dig = _digest()
    if (not force and all(os.path.isfile(v[1]) for v in VARIANTS) and os.path.isfile(stamp)
            and open(stamp).read().strip() == dig):
        return LIB
    procs = []
    objs = {}
    for name, lib, defs in VARIANTS:
        odir = os.path.join(LIB_DIR, "obj_" + name)
        os.makedirs(odir, exist_ok=True)
        objs[name] = []
        for src in _sources():
            obj = os.path.join(odir, os.path.basename(src)[:-3] + ".o")
            objs[name].append(obj)
            cmd = [NVCC] + FLAGS + defs + ["-c", src, "-o", obj]
            procs.append((name, src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for name, src, p in procs:
        out, _ = p.communicate()
        log.append("== [%s] %s\n%s" % (name, os.path.basename(src), out))
        if p.returncode != 0:
            failed = True
    with open(os.path.join(LIB_DIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see %s" % os.path.join(LIB_DIR, "build.log"))
    for name, lib, _ in VARIANTS:
        subprocess.check_call([NVCC, "-shared", "-o", lib] + objs[name] + ["-gencode", "arch=compute_100a,code=sm_100a"])
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
