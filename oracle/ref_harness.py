"""Container-only harness that imports the UNMODIFIED reference (yangbang18/CARE) from
/root/reference and runs its own inference path on CPU.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (care_b200/) may import this
module.  It exists to (a) generate the golden fixtures under tests/golden/ (see
oracle/make_golden.py) and (b) validate oracle/care_oracle.py against the real reference
whenever /root/reference is present.  /root/reference does not exist on the GPU box, so every
user of this module must gate on `reference_available()`.

The recipe follows SURVEY.md §8(c): `pytorch_lightning` and `pycocoevalcap` are absent from
this image, so two families of `sys.modules` stubs are registered before the import.  The
reference's arithmetic (models/*, misc/Decoding/*) is executed untouched.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CARE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "Translator.py"))


def _install_stubs():
    import torch.nn as nn

    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        pl.Trainer = object
        sys.modules["pytorch_lightning"] = pl
    names = [
        "pycocoevalcap", "pycocoevalcap.bleu", "pycocoevalcap.bleu.bleu", "pycocoevalcap.rouge",
        "pycocoevalcap.rouge.rouge", "pycocoevalcap.cider", "pycocoevalcap.cider.cider",
        "pycocoevalcap.meteor", "pycocoevalcap.meteor.meteor", "pycocoevalcap.tokenizer",
        "pycocoevalcap.tokenizer.ptbtokenizer",
    ]
    for n in names:
        if n not in sys.modules:
            m = types.ModuleType(n)
            for attr in ("Bleu", "Rouge", "Cider", "Meteor", "PTBTokenizer"):
                setattr(m, attr, object)
            sys.modules[n] = m


_cached = None


def load_reference():
    """Returns (get_framework, get_translator, Constants) from the reference tree."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from models.Framework import get_framework  # noqa
    from models.Translator import get_translator  # noqa
    from config import Constants  # noqa
    _cached = (get_framework, get_translator, Constants)
    return _cached


def build_reference_model(opt, seed=0):
    """Random-init reference model exactly as `Seq2SeqBase._init_weights` leaves it
    (models/Framework.py:115-134), in eval mode."""
    import torch
    get_framework, _, _ = load_reference()
    torch.manual_seed(seed)
    model = get_framework(dict(opt)).eval()
    return model


def run_reference_translate(model, opt, feats):
    import torch
    _, get_translator, _ = load_reference()
    translator = get_translator(dict(opt))
    with torch.no_grad():
        vocab = {i: str(i) for i in range(opt["vocab_size"])}
        return translator.translate_batch([model], {"feats": feats}, vocab=vocab)
