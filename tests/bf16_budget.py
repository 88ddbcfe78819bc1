"""bf16 error budget of the decode path, on the CPU (no GPU needed).

TEST INFRASTRUCTURE (imports the oracle).  Runs the oracle's beam search with bf16 rounding injected at
one storage site at a time - the sites where the CUDA path in bf16 mode stores or feeds a bf16 value - and
reports, against the unrounded fp32 oracle on the same weights and features:
  * max relative logit error and max scaled log-prob error over the teacher-forced steps
    (the prefixes are the fp32 oracle's own, so errors do not compound through diverging beams);
  * how many captions of a free-running beam search stay identical.

    python -m tests.bf16_budget [cfg4] [n_videos] [preset] [w_mode] [sites | fine | designs | fp16sites]

Sites
  w      GEMM weight matrices (every nn.Linear weight; embeddings / LayerNorm / biases stay fp32)
  feat   input features (the encoder GEMM operand; changes the concept ranking)
  mem    the encoder memory [B, Lm, d] (operand of the cross K/V projection)
  x      the residual stream x0..x3 (both the GEMM operand and the residual input of the next LayerNorm)
  xop    only the GEMM-operand copy of x0..x3 (the residual path keeps fp32)
  cache  self-attention q / k / v (the KV cache)
  ckv    cross-attention K / V (projected once per video)
  attn   attention probabilities fed to the PV product, the context vectors and the cross-attention query
  ffn    the FFN hidden activations
"""
import math
import sys
import time

import torch
import torch.nn.functional as F

from oracle import care_oracle as co
from synth.shapes import CONFIGS, make_feats, make_opt
from synth.weights import PRESETS, make_state_dict

PAD = 0


def r16(x, fmt="bf16"):
    if fmt == "bf16":
        return x.bfloat16().float()
    if fmt == "fp16":
        return x.half().float()
    if fmt == "b2":        # bf16 hi + bf16 lo pair (about 16 mantissa bits)
        hi = x.bfloat16().float()
        return hi + (x - hi).bfloat16().float()
    raise ValueError(fmt)


class Emu:
    """Last-position decoder step over the whole prefix with rounding hooks (same math as
    oracle.care_oracle.decoder_hidden restricted to the newest position; CARE / Base tasks)."""

    def __init__(self, sd, opt, sites, wfmt=None, wskip=()):
        # sites: iterable of site names (bf16) or {site: format} with format in bf16 / fp16 / b2
        # wfmt / wskip: round the DECODER-side weight matrices to `wfmt` except those whose name contains an entry of
        # wskip (the encoder / predictor weights stay fp32: the engine multiplies them as split products)
        self.opt = opt
        self.sites = dict(sites) if isinstance(sites, dict) else {s: "bf16" for s in sites}
        self.sd = dict(sd)
        if "w" in self.sites:
            for k, v in sd.items():
                if v.dim() == 2 and "embeddings" not in k and "hybrid_bias" not in k:
                    self.sd[k] = r16(v)
        if wfmt:
            for k, v in sd.items():
                if v.dim() == 2 and "embeddings" not in k and "hybrid_bias" not in k and \
                        not any(x in k for x in tuple(wskip) + ("encoder", "predictor")):
                    self.sd[k] = r16(v, wfmt)

    GROUPS = {"x0": "x", "x1": "x", "x2": "x", "x3": "x", "x0op": "xop", "x1op": "xop", "x2op": "xop", "x3op": "xop",
              "q": "cache", "k": "cache", "v": "cache", "ck": "ckv", "cv": "ckv", "ps": "attn", "cs": "attn",
              "qc": "attn", "pc": "attn", "cc": "attn"}

    def r(self, site, x):
        """Rounds x if the fine site or its group is listed."""
        fmt = self.sites.get(site) or self.sites.get(self.GROUPS.get(site, ""))
        return r16(x, fmt) if fmt else x

    def op(self, i, x):   # GEMM operand i (0: QKV, 1: cross-Q, 2: FFN1, 3: vocabulary) read from the residual stream
        if ("x%d" % i) in self.sites or "x" in self.sites:
            return x      # the stream itself is already rounded
        return self.r("x%dop" % i, x)

    def encode(self, feats):
        sd, opt = self.sd, self.opt
        if "feat" in self.sites:
            feats = [r16(f, self.sites["feat"]) for f in feats]
        enc = co.encoding_phase(sd, opt, feats)
        if "mem" in self.sites:
            enc["encoder_hidden_states"] = r16(enc["encoder_hidden_states"], self.sites["mem"])
        return enc

    def prepare(self, enc):
        sd = self.sd
        mem = enc["encoder_hidden_states"]
        p = "decoder.layers.0.inter_attention.SDPA."
        self.ck = self.r("ck", F.linear(mem, sd[p + "key.weight"], sd[p + "key.bias"]))
        self.cv = self.r("cv", F.linear(mem, sd[p + "value.weight"], sd[p + "value.bias"]))
        self.gsg = enc.get("semantic_hidden_states")

    def _ln(self, name, x):
        return F.layer_norm(x, (x.shape[-1],), self.sd[name + ".weight"], self.sd[name + ".bias"],
                            self.opt["layer_norm_eps"])

    def step_logits(self, ids, vid):
        """ids [R, t] int64 prefixes; vid [R] video index of every row -> logits [R, V] of the last position."""
        sd, opt = self.sd, self.opt
        H = opt["num_attention_heads"]
        R, t = ids.shape
        d = opt["dim_hidden"]
        dh = d // H
        e = "decoder.embedding"
        x = sd[e + ".word_embeddings.weight"][ids] + sd[e + ".position_embeddings.weight"][:t].unsqueeze(0)
        if self.gsg is not None:
            x = x + self.gsg[vid].unsqueeze(1)
        x0_all = self._ln(e + ".LayerNorm", x)
        x0_all = self.r("x0", x0_all)
        a = "decoder.layers.0.intra_attention."
        xin = self.op(0, x0_all)
        k = self.r("k", F.linear(xin, sd[a + "SDPA.key.weight"], sd[a + "SDPA.key.bias"]))
        v = self.r("v", F.linear(xin, sd[a + "SDPA.value.weight"], sd[a + "SDPA.value.bias"]))
        x0 = x0_all[:, -1]
        q = self.r("q", F.linear(self.op(0, x0), sd[a + "SDPA.query.weight"], sd[a + "SDPA.query.bias"]))
        s = torch.einsum("rhd,rthd->rht", q.view(R, H, dh), k.view(R, t, H, dh)) / math.sqrt(dh)
        s = s.masked_fill(ids.eq(PAD).unsqueeze(1), -1e9)
        ctx = self._pv(s, v.view(R, t, H, dh), "ps", "cs").reshape(R, d)
        x1 = self._ln(a + "LayerNorm", F.linear(ctx, sd[a + "dense.weight"], sd[a + "dense.bias"]) + x0)
        x1 = self.r("x1", x1)
        c = "decoder.layers.0.inter_attention."
        qc = self.r("qc", F.linear(self.op(1, x1), sd[c + "SDPA.query.weight"], sd[c + "SDPA.query.bias"]))
        ck, cv = self.ck[vid], self.cv[vid]
        Lm = ck.shape[1]
        s = torch.einsum("rhd,rlhd->rhl", qc.view(R, H, dh), ck.view(R, Lm, H, dh)) / math.sqrt(dh)
        hb = c + "SDPA.hybrid_bias"
        if hb in sd:
            s = s + sd[hb][None]
        ctx = self._pv(s, cv.view(R, Lm, H, dh), "pc", "cc").reshape(R, d)
        x2 = self._ln(c + "LayerNorm", F.linear(ctx, sd[c + "dense.weight"], sd[c + "dense.bias"]) + x1)
        x2 = self.r("x2", x2)
        f = "decoder.layers.0.ffn."
        h = self.r("ffn", torch.relu(F.linear(self.op(2, x2), sd[f + "dense1.weight"], sd[f + "dense1.bias"])))
        x3 = self._ln(f + "LayerNorm", F.linear(h, sd[f + "dense2.weight"], sd[f + "dense2.bias"]) + x2)
        x3 = self.r("x3", x3)
        return F.linear(self.op(3, x3), sd["cls_head.tgt_word_prj.weight"])

    def _pv(self, s, v, p_site, c_site):
        # s [R, H, n]; v [R, n, H, dh]: fp32 softmax statistics, unnormalised bf16 probabilities into the PV
        # product (as an MMA kernel does), division by the fp32 sum afterwards
        m = s.max(dim=-1, keepdim=True)[0]
        p = torch.exp(s - m)
        z = p.sum(dim=-1, keepdim=True)
        ctx = torch.einsum("rhn,rnhd->rhd", self.r(p_site, p), v) / z
        return self.r(c_site, ctx)


def beam_search(emu, opt, feats, record=None, replay=None):
    """Free-running beam search driven by the oracle's own VideoBeam (so the finish rules are the
    reference's).  record: list that receives (ids, vid) per step.  Returns hyps, scores."""
    K, max_len = opt["beam_size"], opt["max_len"]
    with torch.no_grad():
        enc = emu.encode(feats)
        emu.prepare(enc)
        B = feats[0].shape[0]
        beams = [co.VideoBeam(K, max_len, opt.get("topk", 1)) for _ in range(B)]
        active = list(range(B))
        for t in range(1, max_len):
            ids = torch.stack([beams[i].prefixes() for i in active]).view(-1, t)
            vid = torch.tensor(active).repeat_interleave(K)
            if record is not None:
                record.append((ids.clone(), vid.clone()))
            logp = torch.log_softmax(emu.step_logits(ids, vid), dim=1).view(len(active), K, -1)
            active = [i for pos, i in enumerate(active) if not beams[i].advance(logp[pos])]
            if not active:
                break
    hyps, scores = [], []
    for b in beams:
        sc, tk = b.ranked(opt.get("beam_alpha", 1.0))
        scores.append(sc[0])
        hyps.append(b.backtrack(tk[0][1], tk[0][0]))
    return hyps, scores


def teacher_forced_errors(ref, emu, steps, feats):
    """Per-step errors on the fp32 run's own prefixes."""
    with torch.no_grad():
        emu.prepare(emu.encode(feats))
        worst_logit = worst_lp = worst_lp_abs = 0.0
        for ids, vid in steps:
            a = ref.step_logits(ids, vid)
            b = emu.step_logits(ids, vid)
            scale = a.abs().max().item()
            worst_logit = max(worst_logit, (a - b).abs().max().item() / scale)
            la, lb = torch.log_softmax(a, 1), torch.log_softmax(b, 1)
            top = la > -12
            err = ((la - lb).abs() * top).max().item()
            worst_lp_abs = max(worst_lp_abs, err)
            worst_lp = max(worst_lp, err / max(1.0, scale))
    return worst_logit, worst_lp, worst_lp_abs


def distribution_stats(ref, steps):
    """Peakedness of the fp32 model on its own beam prefixes: top-1 probability, entropy (nats)."""
    p1, ent = [], []
    with torch.no_grad():
        for ids, vid in steps:
            lp = torch.log_softmax(ref.step_logits(ids, vid), 1)
            p = lp.exp()
            p1.append(p.max(dim=1)[0])
            ent.append(-(p * lp).sum(1))
    p1, ent = torch.cat(p1), torch.cat(ent)
    return dict(top1_median=p1.median().item(), top1_mean=p1.mean().item(), entropy_mean=ent.mean().item(),
                entropy_median=ent.median().item())


ACT = ("feat", "mem", "x", "cache", "ckv", "attn", "ffn")


def fmt_sites(sites):
    return "+".join(k if v == "bf16" else "%s:%s" % (k, v) for k, v in sites.items()) or "(none)"


def main():
    """argv: config, videos, weight preset, w_mode.  w_mode "fp32": the oracle keeps fp32 weights (the rows
    with `w` round them); "bf16": the weight set itself is bf16-representable (both sides see identical
    weights, as with a checkpoint stored in bf16) - isolates activation rounding."""
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    preset = sys.argv[3] if len(sys.argv) > 3 else "trained"
    w_mode = sys.argv[4] if len(sys.argv) > 4 else "fp32"
    which = sys.argv[5] if len(sys.argv) > 5 else "sites"
    opt = make_opt(**CONFIGS[cfg])
    sd = make_state_dict(opt, **PRESETS[preset])
    if w_mode == "bf16":
        sd = Emu(sd, opt, ("w",)).sd
    feats = make_feats(opt, n, seed=21)
    ref = Emu(sd, opt, ())
    steps = []
    t0 = time.time()
    ref_h, ref_s = beam_search(ref, opt, feats, record=steps)
    lens = sorted(len(h) for h in ref_h)
    print("%s %s weights=%s n=%d  fp32 run %.1fs  caption lengths min/median/max %d/%d/%d  stats %s" % (
        cfg, preset, w_mode, n, time.time() - t0, lens[0], lens[len(lens) // 2], lens[-1],
        {k: round(v, 3) for k, v in distribution_stats(ref, steps[1::4]).items()}))
    oh, _ = co.ar_translate(sd, opt, [f[:4] for f in feats])
    assert [h[0] for h in oh] == ref_h[:4], "emulator without rounding must equal the oracle"
    tf_steps = steps[0::3]
    allb = {s_: "bf16" for s_ in ACT}
    if which == "sites":       # one site at a time, then everything (the round-1 bf16 mode)
        variants = [{s_: "bf16"} for s_ in (("w",) if w_mode == "fp32" else ()) + ACT] + [{"xop": "bf16"}, allb]
        if w_mode == "fp32":
            variants.append(dict(allb, w="bf16"))
    elif which == "fine":      # every operand / stored tensor separately
        variants = [{s_: "bf16"} for s_ in ("x0op", "x1op", "x2op", "x3op", "x0", "x1", "x2", "x3", "q", "k", "v",
                                            "ck", "cv", "ps", "cs", "qc", "pc", "cc", "ffn", "mem")]
    else:                      # candidate designs
        dec16 = {s_: "fp16" for s_ in ("mem", "x", "cache", "ckv", "attn", "ffn")}
        variants = [
            allb,
            dict(allb, feat="b2"),                                   # encoder operand as a hi/lo pair
            {s_: "fp16" for s_ in ACT},                              # fp16 activations
            dict(dec16, feat="b2"),
            dict(dec16),                                             # encoder in fp32
            {k: v for k, v in dict(dec16, xop="fp16").items() if k != "x"},   # + fp32 residual stream
            {s_: "b2" for s_ in ACT},
        ]
        if w_mode == "fp32":
            variants = [dict(v, w="bf16") for v in variants]
    if which == "fp16sites":   # round 2: where the fp16 mode's error comes from, one site at a time
        dec16 = {s_: "fp16" for s_ in ("mem", "x", "cache", "ckv", "attn", "ffn")}
        named = [
            ("decoder weights fp16", {}, "fp16", ()),
            ("decoder weights fp16 except vocabulary", {}, "fp16", ("tgt_word_prj",)),
            ("vocabulary weights fp16 only", {}, "fp16", ("decoder",)),
            ("all activations fp16, weights fp32", dec16, None, ()),
            ("x3 operand (vocabulary GEMM input)", {"x3op": "fp16"}, None, ()),
            ("x0..x2 operands (QKV / cross-Q / FFN1 inputs)", {"x0op": "fp16", "x1op": "fp16", "x2op": "fp16"}, None, ()),
            ("residual stream x0..x3 (operand + residual)", {"x": "fp16"}, None, ()),
            ("KV cache q / k / v", {"cache": "fp16"}, None, ()),
            ("cross K / V", {"ckv": "fp16"}, None, ()),
            ("attention probabilities / contexts / cross-Q", {"attn": "fp16"}, None, ()),
            ("FFN hidden", {"ffn": "fp16"}, None, ()),
            ("encoder memory", {"mem": "fp16"}, None, ()),
            ("ALL (the fp16 mode)", dec16, "fp16", ()),
            ("ALL, vocabulary GEMM exact", {k: v for k, v in dict(dec16, x0="fp16", x1="fp16", x2="fp16").items() if k != "x"},
             "fp16", ("tgt_word_prj",)),
        ]
        print("%-56s %10s %10s %10s %12s" % ("rounded to fp16", "logit rel", "lp scaled", "lp abs", "exact match"))
        for name, sites, wfmt, wskip in named:
            e = teacher_forced_errors(ref, Emu(sd, opt, sites, wfmt, wskip), tf_steps, feats)
            h, _ = beam_search(Emu(sd, opt, sites, wfmt, wskip), opt, feats)
            same = sum(int(a == b) for a, b in zip(h, ref_h))
            print("%-56s %10.2e %10.2e %10.2e %8d/%d" % (name, e[0], e[1], e[2], same, n), flush=True)
        return
    print("%-56s %10s %10s %10s %12s" % ("sites rounded", "logit rel", "lp scaled", "lp abs", "exact match"))
    for sites in variants:
        e = teacher_forced_errors(ref, Emu(sd, opt, sites), tf_steps, feats)
        h, _ = beam_search(Emu(sd, opt, sites), opt, feats)
        same = sum(int(a == b) for a, b in zip(h, ref_h))
        print("%-56s %10.2e %10.2e %10.2e %8d/%d" % (fmt_sites(sites), e[0], e[1], e[2], same, n), flush=True)


if __name__ == "__main__":
    main()
