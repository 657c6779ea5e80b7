"""merian_b200 — B200-native path tracer behind merian-graph's `pt` node (host side).

Only what the hot path needs: the scene tables (scene.py), synthetic inputs (synthetic.py), the ctypes binding of the
C-ABI library (libmpt.so, built from csrc/) and the Python mirror of the reference's node interface (nodes.py).
"""
from . import abi  # noqa: F401
