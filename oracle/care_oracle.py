"""CPU restatement of the reference's caption-decode path (yangbang18/CARE).

TEST INFRASTRUCTURE ONLY — this module is the *checker*.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` legs may import it; the product package
(care_b200/) never does and has no CPU fallback.

What it restates.  The reference is pure Python whose arithmetic lives in PyTorch (an
un-vendored dependency, `requirement.txt:1`, unpinned; this image has torch 2.11.0).  The
restatement therefore keeps PyTorch's CPU fp32 operators (`F.linear`, `F.layer_norm`,
`softmax`, `log_softmax`, `topk`) at the reference's own call sites and re-writes everything
around them — module plumbing, masks, the per-video beam bookkeeping, the mask-predict loop —
as plain functions over a flat `state_dict`.  It deliberately keeps the reference's *structure*
(full-prefix recompute every step, memory K/V re-projected for every beam row, per-video Python
beam objects, physical compaction of finished videos) so that timing it is a fair stand-in for
the reference's CPU path (bench.py `cpu_baseline.kind == "port"`).

Pinning.  The reference ships no tests or golden vectors for this path, so
the oracle is pinned against outputs of the reference itself executed in the build container:
tests/golden/*.json are produced by oracle/make_golden.py (committed) by importing
/root/reference unmodified, and tests/test_oracle_golden.py checks this file against them
(token sequences, concept ids and scores bit-for-bit).  Known answers that are in the
reference's notebooks (18,218,884 parameters and the module tree of MSRVTT CARE base,
notebooks/retrieval_robustness.ipynb:97-187) are checked in tests/test_state_dict_layout.py.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

PAD, UNK, BOS, EOS, MASK, VIS = 0, 1, 2, 3, 4, 5  # config/Constants.py:1-6


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _ln(sd, name, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _heads(x, n_heads):
    # models/components/Attention.py:58-61  [b, l, d] -> [b, h, l, d/h]
    b, l, d = x.shape
    return x.view(b, l, n_heads, d // n_heads).permute(0, 2, 1, 3)


def repeat_rows(x, k):
    # misc/utils.py:244-258 (`enlarge`): each video's row is repeated k times, contiguous
    return x.unsqueeze(1).repeat(1, k, *([1] * (x.dim() - 1))).contiguous().view(x.shape[0] * k, *x.shape[1:])


# --------------------------------------------------------------------------------------
# encoder + predictor  (Framework.encoding_phase)
# --------------------------------------------------------------------------------------
def encode_stream(sd, opt, ch, x):
    """One modality stream.  Embedder: models/Encoder.py:165-168 (Linear, LayerNorm, Dropout=id).
    EncoderWithHighWayBN: models/Encoder.py:184-187, HighWay :210-226, BN1d :229-241 (eval)."""
    p = "encoder.Encoder_%s" % ch.upper()
    h = _lin(sd, p + ".0", x)
    if opt["encoder"] == "Embedder":
        return _ln(sd, p + ".1", h, opt["layer_norm_eps"])
    if opt["encoder"] == "EncoderWithHighWayBN":
        y = torch.tanh(_lin(sd, p + ".1.w1", h))
        gate = torch.sigmoid(_lin(sd, p + ".1.w2", h))
        h = gate * h + (1 - gate) * y
        flat = h.contiguous().view(-1, h.shape[-1])
        flat = F.batch_norm(flat, sd[p + ".2.bn.running_mean"], sd[p + ".2.bn.running_var"],
                            sd[p + ".2.bn.weight"], sd[p + ".2.bn.bias"], False, 0.1, 1e-5)
        return flat.view(*h.shape)
    raise ValueError("encoder %r is outside the hot path" % opt["encoder"])


def merged_concept_probs(scores):
    """models/Predictor/pred_attribute.py:17-46 without a mask: noisy-or over the sequence axis."""
    probs = torch.sigmoid(scores)
    raw = torch.log(torch.clamp(1.0 - probs, 1e-12, 1))
    return 1.0 - torch.exp(raw.sum(dim=1))


def encoding_phase(sd, opt, feats: List[torch.Tensor]) -> Dict[str, torch.Tensor]:
    """models/Framework.py:150-187 with MultipleStreams.forward (models/Encoder.py:85-153),
    Predictor.forward (models/Predictor/base.py:11-15), Predictor_attribute.forward
    (pred_attribute.py:78-131), SemanticContainer.forward (:262-289), NaiveEmbeddings.forward
    (models/components/Embeddings.py:53-87), Predictor_length.forward (pred_length.py:14-22)."""
    modality = opt["modality"]
    feats = feats[:len(modality)]
    hs = [encode_stream(sd, opt, ch, x) for ch, x in zip(modality, feats)]
    means = [h.mean(1) for h in hs]
    m_dec = opt.get("modality_for_decoder") or modality
    m_pred = opt.get("modality_for_predictor") or modality
    enc_dec = torch.cat([h for ch, h in zip(modality, hs) if ch in m_dec], dim=1)
    enc_pred = torch.cat([h for ch, h in zip(modality, hs) if ch in m_pred], dim=1)
    means_pred = [m for ch, m in zip(modality, means) if ch in m_pred]
    out = {"encoder_hidden_states": enc_dec}

    crits = [c for c in opt["crits"] if c != "lang"]
    nets = list(crits) + list(opt.get("predictors_to_be_added", []))
    if opt.get("load_teacher_weights", False) and "length" in nets:
        nets.remove("length")
        nets.append("length")
    eps = opt["layer_norm_eps"]
    for i, kind in enumerate(nets):
        p = "predictor.nets.%d" % i
        if kind == "attribute":
            assert opt.get("attribute_prediction_channel_concat") and opt.get("attribute_prediction_mean_pooling")
            x = torch.cat(means_pred, dim=-1).unsqueeze(1)
            scores = _lin(sd, p + ".prj", x)
            out["preds_attr"] = merged_concept_probs(scores)
            out["avg_prob_attr"] = torch.sigmoid(scores).mean(dim=(1, 2))
        elif kind == "SemanticContainer":
            preds = out["preds_attr"]
            labels = preds.topk(opt["use_attr_topk"], dim=1, sorted=True, largest=True)[1]
            emb = sd[p + ".attr_embs.word_embeddings.weight"][labels]
            emb = emb + sd[p + ".attr_embs.position_embeddings.weight"][: labels.shape[1]].unsqueeze(0)
            out["semantic_embs"] = _ln(sd, p + ".attr_embs.LayerNorm", emb, eps)
            out["semantic_labels"] = labels
            if "emb" in opt.get("use_attr_type", ""):
                out["semantic_hidden_states"] = F.linear(preds, sd[p + ".semantic2hidden.weight"])
        elif kind == "length":
            x = _lin(sd, p + ".net.0", enc_pred.mean(1))
            x = _lin(sd, p + ".net.3", torch.relu(x))
            out["preds_length"] = torch.log_softmax(x, dim=-1)
        else:
            raise ValueError(kind)
    if "concat" in opt.get("use_attr_type", "") and "semantic_embs" in out:
        out["encoder_hidden_states"] = torch.cat((out["encoder_hidden_states"], out["semantic_embs"]), dim=1)
    return out


def decoder_input_keys(opt):
    """models/Framework.py:21-33 restricted to the branches the configs reach."""
    keys = ["encoder_hidden_states"]
    if opt.get("use_attr", False) and "att" in opt.get("use_attr_type", "").lower():
        keys.append("semantic_embs")
    if "emb" in opt.get("use_attr_type", ""):
        keys.append("semantic_hidden_states")
    return keys


# --------------------------------------------------------------------------------------
# decoder  (TransformerDecoder.forward -> DecoderLayer -> MHA -> SDPA -> FFN -> NaiveHead)
# --------------------------------------------------------------------------------------
def _attention(sd, p, opt, q_in, kv_in, mask, context_only=False):
    """MultiHeadAttention.forward (models/components/SubLayers.py:40-81, post-LN) around
    ScaledDotProductAttention.forward (models/components/Attention.py:69-131)."""
    H = opt["num_attention_heads"]
    q = _heads(_lin(sd, p + ".SDPA.query", q_in), H)
    k = _heads(_lin(sd, p + ".SDPA.key", kv_in), H)
    v = _heads(_lin(sd, p + ".SDPA.value", kv_in), H)
    s = torch.matmul(q, k.transpose(-1, -2))
    s = s / math.sqrt(q.shape[-1])
    if mask is not None:
        s = s.masked_fill(mask.unsqueeze(1), -1e9)
    if (p + ".SDPA.hybrid_bias") in sd:
        s = s + sd[p + ".SDPA.hybrid_bias"][None, :, None, :]
    pr = torch.softmax(s, dim=-1)
    ctx = torch.matmul(pr, v).permute(0, 2, 1, 3).contiguous()
    ctx = ctx.view(ctx.shape[0], ctx.shape[1], -1)
    context = _lin(sd, p + ".dense", ctx)
    if context_only:   # has_ln = skip_connection = False (Layers.py:107-108): the caller adds and normalises
        return context
    out = context + q_in
    return _ln(sd, p + ".LayerNorm", out, opt["layer_norm_eps"])


def decoder_hidden(sd, opt, input_ids, inputs, decoding_type=None):
    """models/Decoder/Transformer.py:161-237 (+ Embeddings.forward, models/components/Embeddings.py:134-188;
    DecoderLayer.forward, models/components/Layers.py:157-228; PositionwiseFeedForward.forward,
    SubLayers.py:137-152).  `input_ids` [R, L] int64; returns hidden states [R, L, d]."""
    decoding_type = decoding_type or opt["decoding_type"]
    mem = inputs["encoder_hidden_states"]
    R, L = input_ids.shape
    keypad = input_ids.eq(PAD).unsqueeze(1).expand(-1, L, -1)
    if decoding_type == "NARFormer":
        self_mask = keypad
    else:
        causal = torch.triu(torch.ones((L, L), dtype=torch.uint8), diagonal=1).unsqueeze(0).expand(R, -1, -1)
        self_mask = (keypad + causal).gt(0)
    cross_mask = torch.zeros(R, L, mem.shape[1], dtype=torch.bool)

    e = "decoder.embedding"
    x = sd[e + ".word_embeddings.weight"][input_ids]
    x = x + sd[e + ".position_embeddings.weight"][:L].unsqueeze(0)
    if decoding_type == "NARFormer":
        assert opt["enhance_input"] == 2
        x = x + mem.mean(1).unsqueeze(1).repeat(1, L, 1)
    if "emb" in opt.get("use_attr_type", ""):
        x = x + inputs["semantic_hidden_states"].unsqueeze(1).expand_as(x)
    x = _ln(sd, e + ".LayerNorm", x, opt["layer_norm_eps"])

    lp = "decoder.layers.0"
    x = _attention(sd, lp + ".intra_attention", opt, x, x, self_mask)
    # attr_attention: a second cross-attention over the concept embeddings, no mask
    # (models/components/Layers.py:117-119,140-155,180-225); position per `attr_layer_pos`
    has_attr = opt.get("use_attr", False) and "att" in opt.get("use_attr_type", "att")
    pos = opt.get("attr_layer_pos", "cross2attr")
    if has_attr and pos == "attr2cross":
        x = _attention(sd, lp + ".attr_attention", opt, x, inputs["semantic_embs"], None)
    if has_attr and pos == "parallel":   # Layers.py:188-201
        inter = _attention(sd, lp + ".inter_attention", opt, x, mem, cross_mask, context_only=True)
        attr = _attention(sd, lp + ".attr_attention", opt, x, inputs["semantic_embs"], None, context_only=True)
        x = _ln(sd, lp + ".LayerNorm", x + inter + attr, opt["layer_norm_eps"])
    else:
        x = _attention(sd, lp + ".inter_attention", opt, x, mem, cross_mask)
    if has_attr and pos == "cross2attr":
        x = _attention(sd, lp + ".attr_attention", opt, x, inputs["semantic_embs"], None)
    h = _lin(sd, lp + ".ffn.dense2", torch.relu(_lin(sd, lp + ".ffn.dense1", x)))
    return _ln(sd, lp + ".ffn.LayerNorm", h + x, opt["layer_norm_eps"])


def decoding_phase(sd, opt, input_ids, inputs, last_time_step_logits=False, decoding_type=None):
    """TransformerSeq2Seq.decoding_phase (models/Framework.py:240-269) + NaiveHead (models/Head.py:26-32)."""
    h = decoder_hidden(sd, opt, input_ids, inputs, decoding_type)
    if last_time_step_logits:
        h = h[:, -1, :]
    return F.linear(h, sd["cls_head.tgt_word_prj.weight"])


# --------------------------------------------------------------------------------------
# beam search  (misc/Decoding/Beam.py + Translator_ARFormer)
# --------------------------------------------------------------------------------------
class VideoBeam:
    """Per-video beam state; restates misc/Decoding/Beam.py:4-132."""

    def __init__(self, size, max_len, n_sents=0, bos=BOS, audit=False):
        self.audit = audit
        self.size = size
        self.need = max(size, n_sents)                      # Beam.py:10
        self.max_len = max_len
        self.done = False
        self.scores = torch.zeros(size)
        self.parents: List[torch.Tensor] = []               # prev_ks
        self.tokens = [torch.full((size,), bos, dtype=torch.long)]  # next_ys
        self.finished: List[list] = []
        self.trace: List[dict] = []

    def _record(self, score, t, k):                        # Beam.py:38-43
        self.finished.append([score, t, k])
        return len(self.finished) >= self.need

    def prefixes(self):
        """Beam.py:26-28,112-132: the K prefixes (with BOS), ordered by a descending sort of the scores."""
        order = torch.sort(self.scores, 0, True)[1]
        rows = []
        for k in order.tolist():
            rows.append(self.backtrack(k, len(self.parents), with_bos=True))
        return torch.LongTensor(rows)

    def backtrack(self, k, length, with_bos=False):        # Beam.py:119-132
        hyp = []
        for j in range(length - 1, -1, -1):
            hyp.append(int(self.tokens[j + 1][k]))
            k = int(self.parents[j][k])
        if with_bos:
            hyp.append(int(self.tokens[0][k]))
        return hyp[::-1]

    def advance(self, logp):                                # Beam.py:45-85
        V = logp.shape[1]
        if self.parents:
            cand = logp + self.scores.unsqueeze(1).expand_as(logp)
            last = self.tokens[-1]
            for i in range(last.shape[0]):
                if last[i] == EOS:
                    cand[i] = -1e20
        else:
            cand = logp[0]
        flat = cand.reshape(-1)
        best, idx = flat.topk(self.size, 0, True, True)
        if self.audit and flat.numel() > self.size:   # margin audit only, not part of the algorithm
            wide = flat.topk(self.size + 1, 0, True, True)[0]
            self.trace.append(dict(scores=best.clone(), ids=idx.clone(), runner_up=float(wide[-1])))
        self.scores = best
        parent = idx // V
        self.parents.append(parent)
        self.tokens.append(idx - parent * V)
        new = self.tokens[-1]
        for i in range(new.shape[0]):
            if new[i] == EOS:
                self.done = self._record(float(self.scores[i]), len(self.parents), i)
            if self.done:
                return True
        if len(self.tokens) == self.max_len:
            self.done = True
            if not self.finished:
                for i in range(new.shape[0]):
                    self._record(float(self.scores[i]), len(self.parents), i)
        return self.done

    def ranked(self, alpha):                                # Beam.py:91-101
        for item in self.finished:
            item[0] /= item[1] ** alpha
        self.finished.sort(key=lambda a: -a[0])
        return [s for s, _, _ in self.finished], [(t, k) for _, t, k in self.finished]


def _ar_translate_ensemble(sds, opt, feats):
    K = opt.get("beam_size", 5)
    alpha = opt.get("beam_alpha", 1.0)
    n_best = opt.get("topk", 1)
    max_len = opt.get("max_len", 30)
    with torch.no_grad():
        B = feats[0].shape[0]
        all_inputs = []
        for sd in sds:
            enc = encoding_phase(sd, opt, feats)
            all_inputs.append({k: repeat_rows(enc[k], K) for k in decoder_input_keys(opt)})
        bos = opt.get("ar_token_id") if opt.get("ar_token_id") is not None else BOS
        beams = [VideoBeam(K, max_len, n_best, bos=bos) for _ in range(B)]
        active = list(range(B))
        for t in range(1, max_len):
            ids = torch.stack([beams[i].prefixes() for i in active]).view(-1, t)
            lps = [torch.log_softmax(decoding_phase(sd, opt, ids, inp, last_time_step_logits=True), dim=1)
                   for sd, inp in zip(sds, all_inputs)]
            logp = torch.stack(lps, dim=0).mean(0).view(len(active), K, -1)      # Translator.py:131
            still = [i for pos, i in enumerate(active) if not beams[i].advance(logp[pos])]
            if not still:
                break
            pos_of = {i: p for p, i in enumerate(active)}
            keep = torch.LongTensor([pos_of[i] for i in still])
            all_inputs = [{k: _select_rows(v, keep, len(active), K) for k, v in inp.items()} for inp in all_inputs]
            active = still
    hyps, scores = [], []
    for b in beams:
        sc, tk = b.ranked(alpha)
        n_best = min(n_best, len(sc))
        scores.append(sc[:n_best])
        hyps.append([b.backtrack(k, t) for t, k in tk[:n_best]])
    return hyps, scores


def _select_rows(x, keep, n_prev, k):
    # Translator.collect_active_part (models/Translator.py:191-209)
    rest = x.shape[1:]
    return x.view(n_prev, -1).index_select(0, keep).view(len(keep) * k, *rest)


def ar_translate(sd, opt, feats, return_trace=False):
    """Translator_ARFormer.translate_batch (models/Translator.py:35-85) with beam_decode_step
    (:91-109), predict_word (:111-133), collect_active_* (:135-209) and
    collect_hypothesis_and_scores (:211-220).  `sd` may be a list of state dicts: model ensembling,
    the beams follow the mean of the models' log-probabilities (:39-52,127-131)."""
    if isinstance(sd, (list, tuple)):
        return _ar_translate_ensemble(list(sd), opt, feats)
    K = opt.get("beam_size", 5)
    alpha = opt.get("beam_alpha", 1.0)
    n_best = opt.get("topk", 1)
    max_len = opt.get("max_len", 30)
    with torch.no_grad():
        enc = encoding_phase(sd, opt, feats)
        B = feats[0].shape[0]
        inputs = {k: repeat_rows(enc[k], K) for k in decoder_input_keys(opt)}
        bos = opt.get("ar_token_id") if opt.get("ar_token_id") is not None else BOS   # Translator.py:61
        beams = [VideoBeam(K, max_len, n_best, bos=bos, audit=return_trace) for _ in range(B)]
        active = list(range(B))
        for t in range(1, max_len):
            ids = torch.stack([beams[i].prefixes() for i in active]).view(-1, t)
            logits = decoding_phase(sd, opt, ids, inputs, last_time_step_logits=True)
            logp = torch.log_softmax(logits, dim=1).view(len(active), K, -1)
            still = [i for pos, i in enumerate(active) if not beams[i].advance(logp[pos])]
            if not still:
                break
            pos_of = {i: p for p, i in enumerate(active)}
            keep = torch.LongTensor([pos_of[i] for i in still])
            inputs = {k: _select_rows(v, keep, len(active), K) for k, v in inputs.items()}
            active = still
    hyps, scores = [], []
    for b in beams:
        sc, tk = b.ranked(alpha)
        # models/Translator.py:215: `n_best` is overwritten inside the loop, so a video with fewer
        # finished hypotheses than n_best also truncates every later video (reference quirk, kept).
        n_best = min(n_best, len(sc))
        scores.append(sc[:n_best])
        hyps.append([b.backtrack(k, t) for t, k in tk[:n_best]])
    if return_trace:
        return hyps, scores, dict(enc=enc, beams=beams)
    return hyps, scores


# --------------------------------------------------------------------------------------
# mask-predict  (Translator_NARFormer + na_algorithms.MaskPredict)
# --------------------------------------------------------------------------------------
def length_candidates(opt, enc):
    """Translator_NARFormer.predict_length_beam (models/Translator.py:307-318)."""
    n = opt["length_beam_size"]
    if "preds_length" in enc:
        beam = enc["preds_length"].topk(n, dim=1)[1] + opt.get("length_bias", 0)
        beam[beam < 4] = 4
        beam[beam > opt["max_len"]] = opt["max_len"]
        return beam
    lo, hi = opt.get("na_length_range", [5, 11])
    return torch.arange(lo, hi, dtype=torch.long).unsqueeze(0).repeat(enc["encoder_hidden_states"].shape[0], 1)


def _nar_pass(sd, opt, inputs, tokens, pad_mask):
    """Algorithm_Base.generate_non_autoregressive (misc/Decoding/na_algorithms.py:67-82) with
    generate_step_with_prob (:6-14); eos_mask is empty because no <eos> is ever placed
    (models/Translator.py:270,282-284: add_eos=False)."""
    logits = decoding_phase(sd, opt, tokens, inputs)
    probs = F.softmax(logits, dim=-1)
    max_probs, idx = probs.max(dim=-1)
    idx[pad_mask] = PAD
    max_probs[pad_mask] = 1.0
    return idx, max_probs


def _teacher_probs(teacher, t_inputs, opt, tokens, pad_mask, is_last):
    """Algorithm_Base.scoring_by_teacher (misc/Decoding/na_algorithms.py:92-126): p_teacher(y_t | y_<t) of the
    student's current tokens from one teacher-forced pass of the auto-regressive teacher; ones when there is no
    teacher or the decision flags (masking_decision / no_candidate_decision, :29-32) switch the rescoring off."""
    ones = torch.ones(tokens.shape, dtype=torch.float32)
    if teacher is None:
        return ones
    if is_last and opt.get("no_candidate_decision", False):
        return ones
    if not is_last and not opt.get("masking_decision", False):
        return ones
    mapping = teacher.get("vocab_mapping")
    tok = mapping[tokens] if mapping is not None else tokens
    with_bos = torch.cat([torch.full((tok.shape[0], 1), BOS, dtype=tok.dtype), tok], dim=1)
    logits = decoding_phase(teacher["sd"], teacher["opt"], with_bos[:, :-1], t_inputs)
    probs = F.softmax(logits, dim=-1).gather(2, tok.unsqueeze(2)).squeeze(2)
    probs[pad_mask] = 1.0
    return probs   # (eos_mask is empty: no <eos> is ever placed, Translator.py:270,282-284)


def nar_translate(sd, opt, feats, return_trace=False, teacher=None):
    """Translator_NARFormer.translate_batch (models/Translator.py:240-305) running
    MaskPredict.generate (misc/Decoding/na_algorithms.py:152-197), select_worst (:128-137).  `teacher`
    (dict sd / opt / vocab_mapping, optional): the auto-regressive model that rescores the candidates
    (Translator.py:250-264, na_algorithms.py:92-126)."""
    with torch.no_grad():
        enc = encoding_phase(sd, opt, feats)
        B = feats[0].shape[0]
        beam = length_candidates(opt, enc)
        n_len = beam.shape[1]
        inputs = {k: repeat_rows(enc[k], n_len) for k in decoder_input_keys(opt)}
        t_inputs = None
        if teacher is not None:
            t_enc = encoding_phase(teacher["sd"], teacher["opt"], feats)
            t_inputs = {k: repeat_rows(t_enc[k], n_len) for k in decoder_input_keys(teacher["opt"])}
        Lmax = int(beam.max())
        tri = torch.triu(torch.ones(Lmax, Lmax, dtype=torch.long), 1)
        length_mask = torch.stack([tri[beam[b] - 1] for b in range(B)], dim=0)
        tokens = (1 - length_mask) * MASK + length_mask * PAD
        tokens = tokens.view(B * n_len, Lmax)

        pad_mask = tokens.eq(PAD)
        seq_lens = Lmax - pad_mask.sum(dim=1)
        use_ct = opt.get("use_ct", False)
        T = opt.get("iterations", 5) + (1 if use_ct else 0)
        if use_ct:   # get_coarse_grained_templates, na_algorithms.py:60-65
            tokens[tokens.eq(MASK)] = VIS
            tokens, probs = _nar_pass(sd, opt, inputs, tokens, pad_mask)
            probs[tokens.eq(MASK)] = 0.0
        else:
            tokens, probs = _nar_pass(sd, opt, inputs, tokens, pad_mask)
        for c in range(1, T):
            corr = _teacher_probs(teacher, t_inputs, opt, tokens, pad_mask, is_last=False)
            if use_ct and c == 1:
                mask_ind = tokens == MASK
            else:
                num_mask = (seq_lens.float() * (1.0 - (c / T))).long()
                mask_ind = torch.zeros_like(probs)
                worst = probs * corr
                for i in range(mask_ind.shape[0]):
                    ind = worst[i].topk(max(1, int(num_mask[i])), largest=False, sorted=False)[1]
                    mask_ind[i, ind] = 1
                mask_ind = mask_ind.bool()
            tokens[mask_ind] = MASK
            new_tokens, new_probs = _nar_pass(sd, opt, inputs, tokens, pad_mask)
            tokens[mask_ind] = new_tokens[mask_ind]
            probs[mask_ind] = new_probs[mask_ind]
        lprobs = (probs * _teacher_probs(teacher, t_inputs, opt, tokens, pad_mask, is_last=True)).log()

        hyp = tokens.view(B, n_len, Lmax)
        lprobs = lprobs.view(B, n_len, Lmax)
        tgt_len = (1 - length_mask).sum(-1).view(B, n_len)
        avg = lprobs.sum(-1) / (tgt_len.float() ** opt.get("beam_alpha", 1.0))
        best = avg.max(-1)[1]
        g = best.unsqueeze(1).unsqueeze(2).repeat(1, 1, Lmax)
        out_h = hyp.gather(1, g).tolist()
        out_p = lprobs.gather(1, g).tolist()
    if return_trace:
        return out_h, out_p, dict(enc=enc, beam=beam, avg=avg, all_tokens=hyp, all_lprobs=lprobs)
    return out_h, out_p


def translate(sd, opt, feats):
    if opt["decoding_type"] == "NARFormer":
        return nar_translate(sd, opt, feats)
    return ar_translate(sd, opt, feats)
