"""`Framework` plugin: the reference's `get_framework(opt)` object (models/Framework.py) backed by
the CUDA engine.

Same constructor input (`opt`), same `state_dict` keys (care_b200/layout.py), same three methods the
Translator calls - `encoding_phase`, `prepare_inputs_for_decoder`, `decoding_phase` - plus
`feedforward_step` and `get_keys_to_device`.  Parameters live in ordinary `nn.Parameter`s so
`load_state_dict` / `.to(device)` / `.parameters()` behave as on the reference module; the engine
re-reads them whenever they change.
"""
import os
from typing import Any, Dict, List

import torch
import torch.nn as nn

from . import layout
from .engine import CareEngine


def get_framework(opt: Dict[str, Any]) -> nn.Module:
    """reference: models/Framework.py:14-51"""
    if "rnn" in opt["decoder"].lower():
        raise ValueError("RNN decoders are outside the accelerated hot path")
    return TransformerSeq2Seq(opt)


class _Node(nn.Module):
    """Bare container used to reproduce the reference's dotted parameter names."""


def _register(root: nn.Module, dotted: str, tensor: torch.Tensor, buffer: bool):
    parts = dotted.split(".")
    node = root
    for p in parts[:-1]:
        if p not in node._modules:
            node.add_module(p, _Node())
        node = node._modules[p]
    if buffer:
        node.register_buffer(parts[-1], tensor)
    else:
        node.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class TransformerSeq2Seq(nn.Module):
    def __init__(self, opt: Dict[str, Any], init_weights: bool = True):
        super().__init__()
        self.opt = dict(opt)
        if opt.get("decoder") not in ("TransformerDecoder", "TwoStageTransformerDecoder"):
            raise ValueError("decoder %r is outside the accelerated hot path" % opt.get("decoder"))
        if opt.get("cls_head", "NaiveHead") != "NaiveHead":
            raise ValueError("cls_head %r is outside the accelerated hot path" % opt.get("cls_head"))
        if opt.get("num_hidden_layers_decoder", 1) != 1:
            raise ValueError("only the 1-layer decoder of config/archs.yaml is accelerated")
        for flag in ("with_category", "pointer", "RPE", "transformer_pre_ln", "compositional_intra",
                     "compositional_inter", "compositional_ffn", "pretrained_embs_path"):
            if opt.get(flag):
                raise ValueError("option %r is outside the accelerated hot path" % flag)
        if opt.get("use_attr", False) and opt.get("use_attr_type", "") not in ("emb_concat", "_att", "emb_att", "_concat"):
            raise ValueError("use_attr_type %r is outside the accelerated hot path (CARE: G1Lc = emb_concat, "
                             "CABase: G0L1 = _att)" % opt.get("use_attr_type"))
        if layout.has_attr_attention(opt):
            if opt.get("attr_layer_pos", "cross2attr") not in ("cross2attr", "attr2cross", "parallel"):
                raise ValueError("attr_layer_pos %r is outside the accelerated hot path" % opt.get("attr_layer_pos"))
            if opt.get("add_hybrid_attention_bias", False):
                raise ValueError("attr_attention with a hybrid attention bias is not a valid reference configuration")
        elif opt.get("attr_layer_pos", "cross2attr") == "parallel":
            # Layers.py:107-108: the video cross-attention would lose its residual and LayerNorm with nothing to
            # replace them (the parallel merge only exists next to an attr_attention)
            raise ValueError("attr_layer_pos 'parallel' needs the attr_attention layer (use_attr_type '..att')")
        if not opt.get("trainable_pe", False):
            raise ValueError("sinusoidal position embeddings are outside the accelerated hot path")
        # options the engine would silently compute differently from the reference: refuse them loudly
        if opt.get("fusion", "temporal_concat") != "temporal_concat":   # Encoder.py:125-153
            raise ValueError("fusion %r is outside the accelerated hot path (only temporal_concat)" % opt.get("fusion"))
        if opt.get("decoding_type") == "NARFormer" and opt.get("enhance_input", 2) != 2:   # Transformer.py:182-189
            raise ValueError("enhance_input %r is outside the accelerated hot path (NARFormer adds the memory mean: 2)"
                             % opt.get("enhance_input"))
        if opt.get("decoding_type", "ARFormer") not in ("ARFormer", "NARFormer"):
            raise ValueError("decoding_type %r is outside the accelerated hot path" % opt.get("decoding_type"))
        if opt.get("position_embeddings_na") or opt.get("with_bn_embedding"):
            raise ValueError("position_embeddings_na / with_bn_embedding are outside the accelerated hot path")
        self.backbone = None  # translate.py:213 reads `.captioner.backbone`
        for name, (shape, kind) in layout.param_specs(opt).items():
            dtype = torch.long if kind == "bn_count" else torch.float32
            _register(self, name, torch.zeros(shape, dtype=dtype), kind in layout.BUFFER_KINDS)
        # reference: models/Framework.py:21-33
        self.input_keys_for_decoder = ["encoder_hidden_states"]
        if opt.get("use_attr", False) and "att" in opt.get("use_attr_type", "").lower():
            self.input_keys_for_decoder.append("semantic_embs")
        if "emb" in opt.get("use_attr_type", ""):
            self.input_keys_for_decoder.append("semantic_hidden_states")
        self._kinds = {n: k for n, (_, k) in layout.param_specs(opt).items()}
        self._engine = None
        self.precision = opt.get("care_precision") or os.environ.get("CARE_B200_PRECISION", "fp16")
        if self.precision not in ("fp32", "fp16", "bf16"):
            raise ValueError("care_precision must be 'fp32', 'fp16' or 'bf16' (got %r)" % self.precision)
        if init_weights:
            self._init_weights()

    # -- weights ------------------------------------------------------------------------------
    def _init_weights(self):
        """Same distributions as the reference's `_init_weights` (models/Framework.py:115-134)."""
        tensors = dict(self.named_parameters())
        tensors.update(dict(self.named_buffers()))
        for name, t in tensors.items():
            kind = self._kinds[name]
            with torch.no_grad():
                if kind in ("linear_w", "emb", "emb_pad"):
                    nn.init.xavier_uniform_(t)
                    if kind == "emb_pad":
                        t[layout.PAD].zero_()
                elif kind in ("ln_w", "bn_w", "bn_var"):
                    t.fill_(1.0)
                else:
                    t.zero_()
        self._engine = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        """`strict=False` tolerates MISSING keys as on the reference module, but a key this model does not have
        would be a module the engine never runs (the reference would run it): that is fatal here."""
        own = set(self.state_dict().keys())
        extra = [k for k in state_dict.keys() if k not in own]
        if extra:
            raise ValueError("state_dict holds %d tensor(s) outside the accelerated hot path, e.g. %s - the checkpoint's "
                             "opt selects modules this engine does not compute" % (len(extra), extra[:4]))
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._engine = None
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._engine = None
        return out

    def set_precision(self, precision: str):
        self.precision = precision
        self._engine = None
        return self

    def engine(self) -> CareEngine:
        if self._engine is None:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("care_b200 has no CPU path: move the model to a CUDA device first")
            self._engine = CareEngine(self.opt, self.state_dict(), dev, self.precision)
        return self._engine

    # -- reference API ------------------------------------------------------------------------
    def get_keys_to_device(self, teacher_forcing=False, **kwargs):
        # reference: models/Framework.py:136-148
        keys = ["feats", "input_ids"]
        for k in self.input_keys_for_decoder:
            if "hidden_states" not in k:
                keys.append(k)
        return keys

    def encoding_phase(self, feats: List[torch.Tensor], **kwargs) -> Dict[str, torch.Tensor]:
        """reference: models/Framework.py:150-187.  Returns the same keys; `encoder_hidden_states` already
        holds the concept embeddings in its last `use_attr_topk` rows (the "concat")."""
        eng = self.engine()
        feats = [f.to(eng.device, non_blocking=True) for f in feats]
        with torch.no_grad():
            out = eng.encode(feats)
        if eng.concat_concepts and "semantic_embs" not in out:
            out["semantic_embs"] = out["encoder_hidden_states"][:, eng.enc_len:, :]
        if "preds_length_logits" in out:   # pred_length.py:22 (API decoration; the decode ranks the logits)
            out["preds_length"] = torch.log_softmax(out["preds_length_logits"][:, :eng.max_len], dim=-1)
        return out

    def prepare_inputs_for_decoder(self, encoding_phase_outputs, batch):
        # reference: models/Framework.py:189-204
        inputs = {}
        for key in self.input_keys_for_decoder:
            if key in encoding_phase_outputs:
                inputs[key] = encoding_phase_outputs[key]
            elif key in batch:
                inputs[key] = batch[key]
            else:
                raise KeyError("the input key `%s` can not be found in `encoding_phase_outputs` %s nor `batch` %s"
                               % (key, list(encoding_phase_outputs.keys()), list(batch.keys())))
        return inputs

    def decoding_phase(self, input_ids, inputs_for_decoder, last_time_step_logits=False, **kwargs):
        """reference: models/Framework.py:240-269 (stateless, full prefix).  Rows of `inputs_for_decoder`
        may be per video or already repeated per beam; the engine only needs rows_per_video."""
        eng = self.engine()
        with torch.no_grad():
            logits = eng.sequence_logits(input_ids, inputs_for_decoder, last_only=last_time_step_logits,
                                         decoding_type=kwargs.get("decoding_type"))
        return {"logits": logits}

    def feedforward_step(self, batch, **kwargs):
        # reference: models/Framework.py:215-234 (teacher-forced pass used by the non-latency eval branch)
        enc = self.encoding_phase(batch["feats"], **kwargs)
        inputs = self.prepare_inputs_for_decoder(enc, batch)
        dec = self.decoding_phase(batch["input_ids"], inputs, **kwargs)
        return {**enc, **dec, "schedule_sampling_prob": 0}

    def forward(self, batch, **kwargs):
        return self.feedforward_step(batch, **kwargs)
