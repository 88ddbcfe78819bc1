"""How often the bf16 mode returns exactly the fp32 mode's caption, and how its per-caption score differs
(plain benchmark weights: near-uniform distributions with exact fp32 ties)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import care_b200  # noqa: E402
from synth.shapes import CONFIGS, make_feats, make_opt  # noqa: E402
from synth.weights import SHARP, make_state_dict  # noqa: E402

for cfg, kw, name in (("cfg4", dict(seed=0), "plain"), ("cfg4", dict(seed=5, perturb=True, sharpen=SHARP), "sharp")):
    opt = make_opt(**CONFIGS[cfg])
    sd = make_state_dict(opt, **kw)
    feats = [f.cuda() for f in make_feats(opt, 512, seed=21)]
    res = {}
    for prec in ("fp32", "bf16"):
        m = care_b200.get_framework(dict(opt, care_precision=prec))
        m.load_state_dict(sd)
        m = m.eval().cuda()
        res[prec] = care_b200.get_translator(opt).translate_batch([m], {"feats": feats})
        del m
    same = sum(int(a == b) for a, b in zip(res["fp32"][0], res["bf16"][0]))
    ds = [abs(a[0] - b[0]) for a, b in zip(res["fp32"][1], res["bf16"][1])]
    tok_same = tok_all = 0
    for a, b in zip(res["fp32"][0], res["bf16"][0]):
        n = min(len(a[0]), len(b[0]))
        tok_all += max(len(a[0]), len(b[0]))
        tok_same += sum(int(x == y) for x, y in zip(a[0][:n], b[0][:n]))
    print("%s %s: %d/512 captions identical, %.1f%% tokens identical, score |diff| median %.2e max %.2e" % (
        cfg, name, same, 100.0 * tok_same / tok_all, sorted(ds)[len(ds) // 2], max(ds)))
