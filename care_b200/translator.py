"""`Translator` plugins with the reference's names and signatures (models/Translator.py).

`get_translator(opt)` resolves `Translator_{decoding_type}`; `translate_batch(models, batch, vocab=,
teacher_model_wrapper=)` returns `(hyps, scores)` in the reference's formats.  The whole decode runs
on the device: beam state never visits the host until the final ids come back.
"""
import torch

from .engine import hyps_from_device

__all__ = ("Translator_ARFormer", "Translator_NARFormer", "get_translator")


def get_translator(opt: dict):
    # reference: models/Translator.py:14-19
    class_name = "Translator_{}".format(opt["decoding_type"])
    if class_name not in globals():
        raise ValueError("We can not find the class `{}` in {}".format(class_name, __file__))
    return globals()[class_name](opt)


def _single_model(models):
    if isinstance(models, (list, tuple)):
        if len(models) != 1:
            raise NotImplementedError("model ensembling is outside the accelerated hot path (single model only)")
        return models[0]
    return models


class Translator_ARFormer(object):
    """Beam search driver (reference: models/Translator.py:22-220)."""

    def __init__(self, opt: dict = {}):
        self.beam_size = opt.get("beam_size", 5)
        self.beam_alpha = opt.get("beam_alpha", 1.0)
        self.topk = opt.get("topk", 1)
        self.max_len = opt.get("max_len", 30)
        self.ar_token_id = opt.get("ar_token_id", None)
        if self.ar_token_id is not None:
            raise NotImplementedError("ar_token_id is outside the accelerated hot path")

    def translate_batch(self, models, batch, *args, **kwargs):
        model = _single_model(models)
        with torch.no_grad():
            out = self.decode_on_device(model, batch["feats"])
        return hyps_from_device(*out, self.beam_alpha, self.topk)

    def decode_on_device(self, model, feats, trace=None):
        """Returns device tensors (ids [B, topk, T] int32 PAD-filled, lengths, raw scores, steps)."""
        eng = model.engine()
        if eng.max_len != self.max_len:
            raise ValueError("translator max_len %d != model max_len %d" % (self.max_len, eng.max_len))
        enc = model.encoding_phase(feats)
        B = enc["encoder_hidden_states"].shape[0]
        return eng.ar_decode(enc, B, beam_size=self.beam_size, topk=self.topk, beam_alpha=self.beam_alpha,
                             trace=trace)


class Translator_NARFormer(object):
    """Length-beam + mask-predict driver (reference: models/Translator.py:223-318)."""

    def __init__(self, opt: dict = {}):
        self.opt = opt
        self.paradigm = opt.get("paradigm", "mp")
        if self.paradigm != "mp":
            raise NotImplementedError("only the mask-predict paradigm is on the accelerated hot path")
        self.max_len = opt["max_len"]
        self.length_beam_size = opt["length_beam_size"]
        self.beam_alpha = opt.get("beam_alpha", 1.0)
        self.length_bias = opt.get("length_bias", 0)

    def translate_batch(self, models, batch, teacher_model_wrapper=None, vocab=None):
        if teacher_model_wrapper is not None:
            raise NotImplementedError("teacher rescoring is outside the accelerated hot path")
        model = _single_model(models)
        eng = model.engine()
        with torch.no_grad():
            enc = model.encoding_phase(batch["feats"])
            tokens, lprobs = eng.mask_predict(enc, self.opt, self.length_beam_size, self.length_bias,
                                              self.beam_alpha)
        return tokens.cpu().tolist(), lprobs.cpu().tolist()
