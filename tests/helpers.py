"""Shared helpers for the parity tests: scene setup on both arms and comparison metrics."""
import numpy as np

from merian_b200 import abi, synthetic
from merian_b200.scene import SceneTables, upload_scene

_cache = {}


def tables(name, **kw):
    key = (name, tuple(sorted(kw.items())))
    if key not in _cache:
        _cache[key] = SceneTables(getattr(synthetic, name)(**kw))
    return _cache[key]


def make_oracle(t):
    from oracle.oracle_py import OracleContext
    o = OracleContext()
    upload_scene(o, t)
    return o


def render_params(w, h, spp, L=5, frame=0, flags=abi.RENDER_EMISSION_ON_PRIMARY, offset=0, stride=1, mask=0xFF):
    return abi.RenderParams(w, h, frame, spp, L, mask, flags, offset, stride, 0.0)


def luminance(img):
    return img[..., 0] * 0.299 + img[..., 1] * 0.587 + img[..., 2] * 0.114


def compare_hits(a, b):
    """Per-pixel id comparison of two PackedHitInfo images. Returns (mismatch fraction, max bary diff on matches)."""
    same = ((a["instance_id"] == b["instance_id"]) & (a["primitive_id"] == b["primitive_id"]) &
            (a["geometry_index"] == b["geometry_index"]) & (a["instance_index"] == b["instance_index"]))
    ba = a["barycentrics"].astype(np.float32)
    bb = b["barycentrics"].astype(np.float32)
    d = np.abs(ba - bb).max(axis=-1)
    return 1.0 - same.mean(), float(d[same].max()) if same.any() else 0.0, same


def edge_ties(a, b, same, eps=2e-3):
    """Enumerate mismatching pixels and classify them: a shared-edge tie has BOTH hits on a triangle edge
    (min barycentric weight ~ 0). Returns (n_mismatch, n_ties)."""
    def on_edge(h):
        bb = h["barycentrics"].astype(np.float32)
        w = np.stack([bb[..., 0], bb[..., 1], 1.0 - bb[..., 0] - bb[..., 1]], -1)
        return (np.abs(w).min(axis=-1) <= eps) & (h["instance_id"] != 0xFFFFFFFF)
    mis = ~same
    ties = mis & on_edge(a) & on_edge(b)
    return int(mis.sum()), int(ties.sum())


def rel_mse(img, ref, eps=1e-2):
    l, r = luminance(img), luminance(ref)
    return float(np.mean((l - r) ** 2 / (r ** 2 + eps)))
