"""TEST INFRASTRUCTURE ONLY — the synthetic weight / frame generators the goldens were made with.

They live in ``transformerupscaler_b200/synth.py`` (plain seeded generators, shared with bench.py so that the product's
benchmark does not import ``oracle/``); this module re-exports them under their historical names.
"""
from transformerupscaler_b200.synth import synth_frames, synth_state_dict, _rel_index, _spec  # noqa: F401
