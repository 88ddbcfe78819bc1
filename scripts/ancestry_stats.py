"""How many distinct KV-cache slots per position the K beams of a video actually reference at the last
step (the self-attention kernel currently streams all K slots of every position)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import care_b200  # noqa: E402
from synth.shapes import CONFIGS, make_feats, make_opt  # noqa: E402
from synth.weights import SHARP, make_state_dict  # noqa: E402

for name, kw in (("plain", dict(seed=0)), ("sharp", dict(seed=1, perturb=True, sharpen=SHARP))):
    opt = make_opt(**CONFIGS["cfg4"])
    sd = make_state_dict(opt, **kw)
    model = care_b200.get_framework(dict(opt, care_precision="fp16", care_cuda_graph=False))
    model.load_state_dict(sd)
    model = model.eval().cuda()
    tr = care_b200.get_translator(opt)
    feats = [f.cuda() for f in make_feats(opt, 256, seed=3)]
    eng = model.engine()
    trace = []
    enc = model.encoding_phase(feats)
    eng.ar_decode(enc, 256, beam_size=5, topk=1, trace=trace, early_exit_every=0)
    for step in (5, 10, 15, 20, 29):
        rec = trace[step - 1]
        anc = rec["pre"]["anc"][:, :, :step - 1].long()       # [B, K, step-1]
        live = rec["pre"]["done"] == 0
        if step < 2 or not live.any():
            continue
        onehot = torch.zeros(anc.shape[0], 5, anc.shape[2])
        onehot.scatter_(1, anc, 1.0)
        distinct = onehot.sum(1)[live]                          # [B_live, step-1]
        print("%s step %2d: live videos %3d, mean distinct slots per position %.2f (of 5); last-5-positions %.2f" % (
            name, step, int(live.sum()), distinct.mean().item(), distinct[:, -5:].mean().item()))
    del model, eng
