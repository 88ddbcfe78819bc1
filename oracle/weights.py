"""Moved to synth/weights.py (pure data builders; bench.py's GPU arm imports nothing from oracle/)."""
from synth.weights import *  # noqa: F401,F403
from synth.weights import PRESETS, SHARP, TRAINED, hybrid_length, make_state_dict, param_count  # noqa: F401
