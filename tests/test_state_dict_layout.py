"""Known answers from the reference's notebooks: 18,218,884 parameters and the module tree of
MSRVTT CARE base (notebooks/retrieval_robustness.ipynb:97-187), plus the key/shape layout of the
reference's own state_dict captured in tests/golden/state_dict_layout.json."""
import json
import os

from synth.shapes import CONFIGS, make_opt
from synth.weights import make_state_dict, param_count
from tests.helpers import GOLDEN


def _layout():
    with open(os.path.join(GOLDEN, "state_dict_layout.json")) as f:
        return json.load(f)


def test_param_count_known_answer():
    sd = make_state_dict(make_opt(**CONFIGS["cfg2"]))
    assert param_count(sd) == 18218884


def test_keys_and_shapes_match_reference():
    layout = _layout()
    for cfg, rec in layout.items():
        sd = make_state_dict(make_opt(**CONFIGS[cfg]))
        assert list(sd.keys()) == list(rec["keys"].keys()), cfg
        for k, shape in rec["keys"].items():
            assert list(sd[k].shape) == shape, (cfg, k)
        assert param_count(sd) == rec["n_params"]
