"""Device timeline of Translator.translate_stream at a given batch size: CUDA events at the start and end of every
decode (encode + graph replay) on the decode stream, so that busy time and gaps between consecutive batches are
visible.  Usage (GPU box): python scripts/stream_probe.py [batch]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import care_b200  # noqa: E402
from synth.shapes import CONFIGS, make_feats, make_opt  # noqa: E402
from synth.weights import make_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = 24
opt = make_opt(**CONFIGS["cfg4"])
model = care_b200.get_framework(dict(opt, care_precision="fp16"))
model.load_state_dict(make_state_dict(opt, seed=0))
model = model.eval().cuda()
tr = care_b200.get_translator(opt)
chunks = [make_feats(opt, min(512, B - c), seed=c) for c in range(0, B, 512)]
host = [torch.cat([ch[i] for ch in chunks]).pin_memory() for i in range(len(opt["modality"]))]
dev = [f.cuda() for f in host]
for _ in range(3):
    tr.decode_on_device(model, dev, early_exit_every=0)
torch.cuda.synchronize()

starts, ends, host_t = [], [], []
orig = tr.decode_on_device


def wrapped(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out = orig(*a, **k)
    e1.record()
    host_t.append(time.perf_counter() - t0)
    starts.append(e0)
    ends.append(e1)
    return out


tr.decode_on_device = wrapped
batches = [{"feats": host} for _ in range(N)]
for call in range(3):   # the first call allocates the staging / landing buffers; later calls show the steady cost
    del starts[:], ends[:], host_t[:]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    iter_t = []
    last = t0
    for _ in tr.translate_stream([model], batches):
        now = time.perf_counter()
        iter_t.append(now - last)
        last = now
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print("call %d: %.2f ms per batch; host ms until each of the first 4 results: %s, last 3: %s, after the last yield: %.2f" % (
        call, wall / N * 1e3, ["%.2f" % (t * 1e3) for t in iter_t[:4]], ["%.2f" % (t * 1e3) for t in iter_t[-3:]],
        (t0 + wall - last) * 1e3))
busy = [s.elapsed_time(e) for s, e in zip(starts, ends)]
gaps = [ends[i].elapsed_time(starts[i + 1]) for i in range(N - 1)]
print("batch %d: stream wall %.2f ms per batch -> %.0f captions/s" % (B, wall / N * 1e3, B * N / wall))
print("decode busy (encode + graph) ms: first %s ... median %.3f" % (["%.2f" % b for b in busy[:4]], sorted(busy)[N // 2]))
print("gap between decodes (end i -> start i+1) ms: %s ... median %.3f" % (["%.2f" % g for g in gaps[:6]], sorted(gaps)[len(gaps) // 2]))
print("host time inside decode_on_device ms: median %.3f" % (sorted(host_t)[N // 2] * 1e3))
print("host time per yielded batch ms: median %.3f" % (sorted(iter_t)[len(iter_t) // 2] * 1e3))
# resident loop for comparison
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(N):
    orig(model, dev, early_exit_every=0)
b.record()
torch.cuda.synchronize()
print("resident loop: %.3f ms per batch" % (a.elapsed_time(b) / N))
