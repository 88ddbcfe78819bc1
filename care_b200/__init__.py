"""care_b200 - B200-native implementation of CARE's caption-decode hot path.

Public surface mirrors the reference (yangbang18/CARE): `get_framework`, `get_translator`,
`Model`, `ModelEnsemble`, `load_model`, `load_model_from_arguments`, `get_criterion`.  All compute runs in libcare_b200.so
(hand-written CUDA for sm_100a); importing the package does not need a GPU, running it does.
"""
from .framework import TransformerSeq2Seq, get_framework  # noqa: F401
from .translator import Translator_ARFormer, Translator_NARFormer, get_translator  # noqa: F401
from .criterion import get_criterion  # noqa: F401
from .wrapper import (Model, ModelBase, ModelEnsemble, load_model, load_model_from_arguments,  # noqa: F401
                      modify_opt_if_necessary, to_sentence)

__version__ = "0.1.0"
