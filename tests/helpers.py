"""Shared test helpers: golden loading and case reconstruction."""
import glob
import json
import os

from synth.shapes import CONFIGS, make_feats, make_opt
from synth.weights import make_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLDEN, "*.json"))
                  if os.path.basename(p).startswith(("cfg", "cab")))


def ensemble_golden_names():
    return sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLDEN, "ens*.json")))


def rebuild_ensemble_case(rec):
    opt = make_opt(**{**CONFIGS[rec["config"]], **rec["overrides"]})
    sds = [make_state_dict(opt, **w) for w in rec["weights_list"]]
    feats = make_feats(opt, rec["batch"], seed=rec["feat_seed"])
    return opt, sds, feats


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def rebuild_case(rec, batch=None):
    """(opt, state_dict, feats) exactly as oracle/make_golden.py built them."""
    opt = make_opt(**{**CONFIGS[rec["config"]], **rec["overrides"]})
    sd = make_state_dict(opt, **rec["weights"])
    feats = make_feats(opt, rec["batch"], seed=rec["feat_seed"])
    if batch is not None:
        feats = [f[:batch].contiguous() for f in feats]
    return opt, sd, feats
