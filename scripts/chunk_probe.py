"""Single-call translate_batch over a 4096-video pinned host batch for several pipeline_chunk sizes, and the
device-resident decode time of those chunk sizes."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import care_b200  # noqa: E402
from synth.shapes import CONFIGS, make_feats, make_opt  # noqa: E402
from synth.weights import make_state_dict  # noqa: E402

B = 4096
opt = make_opt(**CONFIGS["cfg4"])
model = care_b200.get_framework(dict(opt, care_precision="fp16"))
model.load_state_dict(make_state_dict(opt, seed=0))
model = model.eval().cuda()
tr = care_b200.get_translator(opt)
chunks = [make_feats(opt, 512, seed=c) for c in range(0, B, 512)]
host = [torch.cat([ch[i] for ch in chunks]).pin_memory() for i in range(len(opt["modality"]))]
dev = [f.cuda() for f in host]

for n in ():
    sub = [f[:n].contiguous() for f in dev]
    for _ in range(3):
        tr.decode_on_device(model, sub, early_exit_every=0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    t0 = time.perf_counter()
    for _ in range(5):
        tr.decode_on_device(model, sub, early_exit_every=0)
    host_ms = (time.perf_counter() - t0) / 5 * 1e3
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("resident decode B=%d: %.2f ms (%.0f captions/s), host enqueue %.2f ms" % (n, ms, n / ms * 1e3, host_ms))

for chunk in (1024, 2048, 3072):
    tr.pipeline_chunk = chunk
    for _ in range(2):
        tr.translate_batch([model], {"feats": host})
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        h, s = tr.translate_batch([model], {"feats": host})
    ms = (time.perf_counter() - t0) / 4 * 1e3
    print("translate_batch(host 4096) chunk=%d: %.2f ms -> %.0f captions/s" % (chunk, ms, B / ms * 1e3))
