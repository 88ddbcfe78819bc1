"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/care_b200.h declares, reports errors without a GPU (no CPU fallback), and the host-side
mirror keeps the reference's checkpoint layout."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from care_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    from care_b200 import _lib
    header = open(os.path.join(ROOT, "include", "care_b200.h")).read()
    declared = set(re.findall(r"\b(care_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "symbol %s declared in the header but not exported" % name
    assert declared == set(_lib.EXPORTED_SYMBOLS)


def test_errors_without_gpu(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.care_ctx_create(ctypes.byref(h), 0)
    assert rc != 0
    assert b"no CPU path" in lib.care_last_error() or b"CUDA" in lib.care_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "care_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py") or f.endswith(".cu") or f.endswith(".cuh"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("care_oracle", "oracle") or "import oracle" not in src
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_framework_layout_and_cpu_refusal():
    import care_b200
    from synth.shapes import CONFIGS, make_opt
    from synth.weights import make_state_dict
    for cfg in ("cfg1", "cfg2", "cfg5", "cab"):
        opt = make_opt(**CONFIGS[cfg])
        m = care_b200.get_framework(opt)
        sd = make_state_dict(opt)
        assert list(m.state_dict().keys()) == list(sd.keys())
        m.load_state_dict(sd, strict=True)
    assert m.backbone is None
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.encoding_phase([torch.zeros(1, 28, 128)])
    tr = care_b200.get_translator(make_opt(**CONFIGS["cfg2"]))
    assert type(tr).__name__ == "Translator_ARFormer"
    tr = care_b200.get_translator(make_opt(**CONFIGS["cfg5"]))
    assert type(tr).__name__ == "Translator_NARFormer"


def test_hyps_from_device_reproduces_nbest_carry_over():
    from care_b200.engine import hyps_from_device
    tok = torch.tensor([[[5, 3, 0], [6, 7, 3]], [[8, 3, 0], [0, 0, 0]], [[9, 9, 3], [4, 3, 0]]], dtype=torch.int32)
    ln = torch.tensor([[2, 3], [2, 0], [3, 2]], dtype=torch.int32)
    sc = torch.tensor([[-1.0, -2.0], [-3.0, 0.0], [-4.0, -5.0]])
    tt = ln.clone()
    hyps, scores = hyps_from_device(tok, ln, sc, tt, 1.0, 2)
    assert hyps == [[[5, 3], [6, 7, 3]], [[8, 3]], [[9, 9, 3]]]      # video 1 truncates video 2 as well
    assert scores[0] == [-0.5, -2.0 / 3] and scores[2] == [-4.0 / 3]
