"""Synthetic `opt` dictionaries and feature generators for the BASELINE.json configs.

SYNTHETIC-DATA INFRASTRUCTURE (shared by the oracle, the tests and bench.py).  The keys are the
ones the reference's inference path reads; values follow
config/archs.yaml:1-26, config/tasks.yaml:10-54, config/methods.yaml:1-59, opts.py:209-210.
"""
import torch

FEAT_DIMS = {"a": 128, "m": 2048, "i": 512, "r": 512}

ARCHS = {
    "base": dict(dim_hidden=512, num_attention_heads=8, intermediate_size=2048),
    "median": dict(dim_hidden=768, num_attention_heads=12, intermediate_size=3072),
    "large": dict(dim_hidden=1024, num_attention_heads=16, intermediate_size=4096),
}


def make_opt(task="CARE", arch="base", vocab_size=10547, beam_size=5, method="Transformer",
             n_frames=28, max_len=30, **over):
    opt = dict(
        encoder="Embedder", decoder="TransformerDecoder", cls_head="NaiveHead",
        decoding_type="ARFormer", fusion="temporal_concat",
        hidden_act="relu", layer_norm_eps=1e-12,
        encoder_dropout_prob=0.5, hidden_dropout_prob=0.5, attention_probs_dropout_prob=0.1,
        trainable_pe=True, num_hidden_layers_decoder=1, num_hidden_layers_encoder=1,
        vocab_size=vocab_size, max_len=max_len, n_frames=n_frames,
        feats="synthetic", enhance_input=2,
        beam_size=beam_size, beam_alpha=1.0, topk=1,
        dim_a=FEAT_DIMS["a"], dim_m=FEAT_DIMS["m"], dim_i=FEAT_DIMS["i"], dim_r=FEAT_DIMS["r"],
    )
    opt.update(ARCHS[arch])
    if task == "Base":
        opt.update(modality="mi", crits=["lang"])
    elif task == "CARE":
        opt.update(
            modality="amir", modality_for_decoder="ami", modality_for_predictor="amir",
            attribute_prediction=True, attribute_prediction_flags="V", attribute_prediction_k=500,
            attribute_prediction_mean_pooling=True, attribute_prediction_channel_concat=True,
            use_attr=True, use_attr_flags="G1Lc", use_attr_type="emb_concat", use_attr_topk=30,
            add_hybrid_attention_bias=True, retrieval_topk=20, retrieval_arch="ViT",
            predictors_to_be_added=["SemanticContainer"], crits=["lang", "attribute"],
        )
    elif task == "CABase":
        # config/tasks.yaml:56-60: no GSG, LSG by a second cross-attention over the concept embeddings
        # ("Cross -> Semantic"), visual-driven concept detection, no hybrid attention bias
        opt.update(
            modality="ami", modality_for_decoder="ami", modality_for_predictor="mi",
            attribute_prediction=True, attribute_prediction_flags="V", attribute_prediction_k=500,
            attribute_prediction_mean_pooling=True, attribute_prediction_channel_concat=True,
            use_attr=True, use_attr_flags="G0L1", use_attr_type="_att", use_attr_topk=30,
            attr_layer_pos="cross2attr", add_hybrid_attention_bias=False, retrieval_topk=20,
            predictors_to_be_added=["SemanticContainer"], crits=["lang", "attribute"],
        )
    else:
        raise ValueError(task)
    if method == "NACF":
        opt.update(
            encoder="EncoderWithHighWayBN", decoder="TwoStageTransformerDecoder",
            decoding_type="NARFormer", length_beam_size=6, iterations=5, beam_alpha=1.35,
            use_ct=True, paradigm="mp", load_teacher_weights=True, length_prediction=True,
            crits=list(opt["crits"]) + ["length"],
        )
    elif method != "Transformer":
        raise ValueError(method)
    opt.update(over)
    return opt


def frames_of(opt, char):
    return opt.get("retrieval_topk", 20) if char == "r" else opt["n_frames"]


def make_feats(opt, bsz, seed=0, device="cpu"):
    """feats ~ N(0,1) fp32, list in opt['modality'] order."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for ch in opt["modality"]:
        out.append(torch.randn(bsz, frames_of(opt, ch), opt["dim_" + ch], generator=g).to(device))
    return out


CONFIGS = {
    # BASELINE.json configs[0..4]
    "cfg1": dict(task="Base", arch="base", vocab_size=9468, beam_size=1),
    "cfg2": dict(task="CARE", arch="base", vocab_size=10547, beam_size=5),
    "cfg3": dict(task="CARE", arch="median", vocab_size=14745, beam_size=5),
    "cfg4": dict(task="CARE", arch="large", vocab_size=14745, beam_size=5),
    "cfg5": dict(task="CARE", arch="base", vocab_size=10547, method="NACF"),
    # the paper's second model (CABase, attr_attention decoder layer)
    "cab": dict(task="CABase", arch="base", vocab_size=10547, beam_size=5),
}
