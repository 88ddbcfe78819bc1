"""Container-only harness that imports the UNMODIFIED reference (yangbang18/CARE) from
/root/reference - or from the archive of its Python packages oracle/_ref/reference_py.zip (oracle/build_ref.py), which is
how it reaches the GPU box - and runs its own inference path on CPU.

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

# zip of the reference's Python packages made by oracle/build_ref.py (imported through zipimport)
_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_py.zip")


def _is_reference(root):
    return bool(root) and (os.path.isfile(os.path.join(root, "models", "Translator.py")) or
                           (root.endswith(".zip") and os.path.isfile(root)))


def _pick_root():
    for root in (os.environ.get("CARE_REFERENCE_ROOT"), "/root/reference", _VENDORED):
        if _is_reference(root):
            return root
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    """True in the build container (/root/reference) and wherever oracle/_ref travelled (the GPU box)."""
    return _is_reference(REFERENCE_ROOT)


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
