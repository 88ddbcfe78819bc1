"""Where the host-fed (e2e) time goes at cfg4 B=4096: H2D alone, decode alone, decode with a concurrent H2D,
host enqueue time of one decode, and the translate_stream loop."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import care_b200  # noqa: E402
from synth.shapes import CONFIGS, make_feats, make_opt  # noqa: E402
from synth.weights import make_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
opt = make_opt(**CONFIGS["cfg4"])
model = care_b200.get_framework(dict(opt, care_precision="fp16"))
model.load_state_dict(make_state_dict(opt, seed=0))
model = model.eval().cuda()
tr = care_b200.get_translator(opt)
chunks = [make_feats(opt, min(512, B - c), seed=c) for c in range(0, B, 512)]
host = [torch.cat([ch[i] for ch in chunks]).pin_memory() for i in range(len(opt["modality"]))]
dev = [f.cuda() for f in host]
stage = [torch.empty_like(f) for f in dev]
nbytes = sum(f.numel() * f.element_size() for f in host)
side = torch.cuda.Stream()


def ev():
    return torch.cuda.Event(enable_timing=True)


for _ in range(3):
    tr.decode_on_device(model, dev, early_exit_every=0)
torch.cuda.synchronize()

# H2D alone
a, b = ev(), ev()
a.record()
for _ in range(3):
    for d, s in zip(stage, host):
        d.copy_(s, non_blocking=True)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print("H2D alone: %.2f ms for %.2f GB -> %.1f GB/s" % (ms, nbytes / 1e9, nbytes / ms / 1e6))

# decode alone: device time and host enqueue time
a, b = ev(), ev()
a.record()
t0 = time.perf_counter()
for _ in range(3):
    tr.decode_on_device(model, dev, early_exit_every=0)
host_ms = (time.perf_counter() - t0) / 3 * 1e3
b.record()
torch.cuda.synchronize()
print("decode alone: %.2f ms device, %.2f ms host enqueue" % (a.elapsed_time(b) / 3, host_ms))

# decode with a concurrent H2D on the side stream
a, b, c0, c1 = ev(), ev(), ev(), ev()
a.record()
with torch.cuda.stream(side):
    c0.record(side)
    for _ in range(3):
        for d, s in zip(stage, host):
            d.copy_(s, non_blocking=True)
    c1.record(side)
for _ in range(3):
    tr.decode_on_device(model, dev, early_exit_every=0)
b.record()
torch.cuda.synchronize()
print("decode with concurrent H2D: %.2f ms decode, %.2f ms per H2D" % (a.elapsed_time(b) / 3, c0.elapsed_time(c1) / 3))

# encoder share
a, b = ev(), ev()
a.record()
with torch.no_grad():
    for _ in range(3):
        model.encoding_phase(dev)
b.record()
torch.cuda.synchronize()
print("encoding_phase alone: %.2f ms" % (a.elapsed_time(b) / 3))

# the stream loop
def run(n):
    got = 0
    for h, s in tr.translate_stream([model], ({"feats": host} for _ in range(n))):
        got += len(h)
    return got


run(2)
torch.cuda.synchronize()
t0 = time.perf_counter()
run(6)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 6 * 1e3
print("translate_stream: %.2f ms per batch -> %.0f captions/s" % (ms, B / ms * 1e3))


# sustained decode (power-capped steady state), and what the bench's instrumentation costs
def sustained(n, every, label):
    a, b = ev(), ev()
    a.record()
    for _ in range(n):
        tr.decode_on_device(model, dev, early_exit_every=every)
    b.record()
    torch.cuda.synchronize()
    print("sustained x%d, early_exit_every=%d %s: %.2f ms per decode" % (n, every, label, a.elapsed_time(b) / n))


sustained(20, 0, "")
sustained(20, 4, "")
import subprocess  # noqa: E402
p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-lms", "50"],
                     stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
time.sleep(0.3)
sustained(20, 4, "+ nvidia-smi -lms 50")
p.terminate()
p.wait()
