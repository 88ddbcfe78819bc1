"""Batch-1 decode (latency mode) a few times; run under `ncu --metrics gpu__time_duration.sum` for a launch list."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import care_b200  # noqa: E402
from synth.shapes import CONFIGS, make_feats, make_opt  # noqa: E402
from synth.weights import make_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
opt = make_opt(**CONFIGS["cfg4"])
model = care_b200.get_framework(dict(opt, care_precision="fp16"))
model.load_state_dict(make_state_dict(opt, seed=0))
model = model.eval().cuda()
tr = care_b200.get_translator(opt)
feats = [f.cuda() for f in make_feats(opt, B, seed=3)]
for _ in range(3):
    out = tr.decode_on_device(model, feats)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = tr.decode_on_device(model, feats)
e1.record()
torch.cuda.synchronize()
print("B=%d: %.3f ms per translate" % (B, e0.elapsed_time(e1) / 5))
