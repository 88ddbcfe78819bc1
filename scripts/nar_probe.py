"""Throughput of the mask-predict (config 5: NACF, 6 length candidates, coarse templates + 5 refinements) path."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import care_b200  # noqa: E402
from synth.shapes import CONFIGS, make_feats, make_opt  # noqa: E402
from synth.weights import make_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
opt = make_opt(**CONFIGS["cfg5"])
model = care_b200.get_framework(dict(opt, care_precision="fp16"))
model.load_state_dict(make_state_dict(opt, seed=0, perturb=True))
model = model.eval().cuda()
tr = care_b200.get_translator(opt)
feats = [f.cuda() for f in make_feats(opt, B, seed=3)]
for _ in range(2):
    out = tr.translate_batch([model], {"feats": feats})
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = tr.translate_batch([model], {"feats": feats})
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("cfg5 NACF mask-predict B=%d: %.2f ms per batch, %.0f captions/s (Lmax=%d)" % (B, ms, B / ms * 1e3, len(out[0][0][0])))
