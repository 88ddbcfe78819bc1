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


def teacher_golden_names():
    return sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLDEN, "nar_*teacher*.json")))


def rebuild_teacher_case(rec):
    """(student opt, student sd, feats, teacher dict for oracle.care_oracle.nar_translate) as
    oracle/make_golden.py::run_teacher_case built them."""
    from oracle.make_golden import permuted_vocab_mapping
    opt = make_opt(**{**CONFIGS["cfg5"], **rec["overrides"]})
    t_opt = make_opt(**CONFIGS["cfg2"])
    sd = make_state_dict(opt, **rec["weights"])
    t_sd = make_state_dict(t_opt, **rec["teacher_weights"])
    feats = make_feats(opt, rec["batch"], seed=rec["feat_seed"])
    mapping = permuted_vocab_mapping(opt["vocab_size"], rec["map_seed"]) if rec["map_seed"] is not None else None
    return opt, sd, feats, dict(sd=t_sd, opt=t_opt, vocab_mapping=mapping)
