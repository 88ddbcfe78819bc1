"""Checkpoint layout of the reference model: parameter / buffer names and shapes per `opt`.

The drop-in contract is the reference's `state_dict` naming, e.g.
`encoder.Encoder_A.0.weight`, `predictor.nets.1.attr_embs.word_embeddings.weight`,
`decoder.layers.0.inter_attention.SDPA.hybrid_bias`, `cls_head.tgt_word_prj.weight`
(module tree printed in the reference's notebooks/retrieval_robustness.ipynb:97-184).
"""
from collections import OrderedDict

PAD = 0


def hybrid_length(opt):
    # reference: models/components/Layers.py:85-90
    modality = opt.get("modality_for_decoder") or opt["modality"]
    n = opt["n_frames"] * len(modality) + opt.get("use_attr_topk", 30)
    if "r" in modality:
        n += opt["retrieval_topk"] - opt["n_frames"]
    return n


def has_attr_attention(opt):
    """reference: models/components/Layers.py:117-119 (`'att' in use_attr_type`, default 'att')."""
    return bool(opt.get("use_attr", False)) and "att" in opt.get("use_attr_type", "att")


def predictor_nets(opt):
    """Order of `predictor.nets.*` (reference: models/Predictor/__init__.py:26-60)."""
    nets = [c for c in opt["crits"] if c != "lang"] + list(opt.get("predictors_to_be_added", []))
    if opt.get("load_teacher_weights", False) and "length" in nets:
        nets.remove("length")
        nets.append("length")
    return nets


def param_specs(opt):
    """OrderedDict name -> (shape, kind); kind in {linear_w, bias, ln_w, ln_b, emb, emb_pad, zeros,
    bn_w, bn_b, bn_mean, bn_var, bn_count}."""
    d = opt["dim_hidden"]
    sp = OrderedDict()

    def linear(name, out_f, in_f, bias=True):
        sp[name + ".weight"] = ((out_f, in_f), "linear_w")
        if bias:
            sp[name + ".bias"] = ((out_f,), "bias")

    def ln(name):
        sp[name + ".weight"] = ((d,), "ln_w")
        sp[name + ".bias"] = ((d,), "ln_b")

    for ch in opt["modality"]:
        p = "encoder.Encoder_%s" % ch.upper()
        linear(p + ".0", d, opt["dim_" + ch])
        if opt["encoder"] == "EncoderWithHighWayBN":
            linear(p + ".1.w1", d, d)
            linear(p + ".1.w2", d, d)
            sp[p + ".2.bn.weight"] = ((d,), "bn_w")
            sp[p + ".2.bn.bias"] = ((d,), "bn_b")
            sp[p + ".2.bn.running_mean"] = ((d,), "bn_mean")
            sp[p + ".2.bn.running_var"] = ((d,), "bn_var")
            sp[p + ".2.bn.num_batches_tracked"] = ((), "bn_count")
        elif opt["encoder"] == "Embedder":
            ln(p + ".1")
        else:
            raise ValueError("encoder %r is outside the accelerated hot path" % opt["encoder"])
    for i, kind in enumerate(predictor_nets(opt)):
        p = "predictor.nets.%d" % i
        if kind == "attribute":
            nm = len(opt.get("modality_for_predictor") or opt["modality"])
            width = d * (nm if opt.get("attribute_prediction_channel_concat", False) else 1)
            linear(p + ".prj", opt["attribute_prediction_k"], width)
        elif kind == "SemanticContainer":
            sp[p + ".attr_embs.word_embeddings.weight"] = ((opt["attribute_prediction_k"], d), "emb")
            sp[p + ".attr_embs.position_embeddings.weight"] = ((opt["use_attr_topk"], d), "emb")
            ln(p + ".attr_embs.LayerNorm")
            if "emb" in opt.get("use_attr_type", ""):
                linear(p + ".semantic2hidden", d, opt["attribute_prediction_k"], bias=False)
        elif kind == "length":
            linear(p + ".net.0", d, d)
            linear(p + ".net.3", opt["max_len"], d)
        else:
            raise ValueError("predictor %r is outside the accelerated hot path" % kind)
    sp["decoder.embedding.word_embeddings.weight"] = ((opt["vocab_size"], d), "emb_pad")
    sp["decoder.embedding.position_embeddings.weight"] = ((opt["max_len"], d), "emb")
    ln("decoder.embedding.LayerNorm")
    L = "decoder.layers.0."
    atts = ["intra_attention", "inter_attention"]
    if has_attr_attention(opt):
        atts.append("attr_attention")   # reference: deepcopy of inter_attention, Layers.py:117-119
    # attr_layer_pos 'parallel' (Layers.py:107-108,121-122): the two cross-attentions have neither a residual nor
    # a LayerNorm of their own; the layer owns one LayerNorm over x + inter_context + attr_context
    parallel = has_attr_attention(opt) and opt.get("attr_layer_pos", "cross2attr") == "parallel"
    for att in atts:
        if att != "intra_attention" and opt.get("add_hybrid_attention_bias", False):
            sp[L + att + ".SDPA.hybrid_bias"] = ((opt["num_attention_heads"], hybrid_length(opt)), "zeros")
        for nm_ in ("query", "key", "value"):
            linear(L + att + ".SDPA." + nm_, d, d)
        linear(L + att + ".dense", d, d)
        if att == "intra_attention" or not parallel:
            ln(L + att + ".LayerNorm")
    if parallel:
        ln(L + "LayerNorm")
    linear(L + "ffn.dense1", opt["intermediate_size"], d)
    linear(L + "ffn.dense2", d, opt["intermediate_size"])
    ln(L + "ffn.LayerNorm")
    linear("cls_head.tgt_word_prj", opt["vocab_size"], d, bias=False)
    return sp


BUFFER_KINDS = ("bn_mean", "bn_var", "bn_count")
