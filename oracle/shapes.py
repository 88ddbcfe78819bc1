"""Moved to synth/shapes.py (pure data builders; bench.py's GPU arm imports nothing from oracle/)."""
from synth.shapes import *  # noqa: F401,F403
from synth.shapes import ARCHS, CONFIGS, FEAT_DIMS, frames_of, make_feats, make_opt  # noqa: F401
