"""Synthetic inputs of the parity tests: the same generators bench.py and the tools use (tools/synth.py)."""
from tools.synth import *  # noqa: F401,F403
from tools.synth import _jitter, _qmul, _rays  # noqa: F401
