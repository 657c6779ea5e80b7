"""the roofline numerator (nodes / triangles per ray) is counted by the CPU oracle walking the SAME BVH8 over the SAME
continuation rays the device traced. These tests pin the pieces bench.py composes: the ray-queue read-back, the agreement of the oracle's
walker with the device on those rays, and the bound device work <= 1.10 x oracle work (the device defers triangle tests, so its hit
distance lags and it may visit a little more; a traversal that wandered would show up here)."""
import numpy as np
import pytest

from merian_b200 import abi
from merian_b200.host_accel import HostAccel
from merian_b200.scene import upload_scene

from helpers import make_oracle, tables

pytestmark = pytest.mark.gpu

TERRAIN = dict(nx=200, ny=100, env_size=(256, 128))


def _frame(gpu_ctx, t, L, w=240, h=136, spp=4, tex1=None):
    import torch
    cam = t.desc.camera.to_abi(0)
    if tex1 is None:
        tex1 = gpu_ctx.primary_visibility_device(cam, w, h)
    P = abi.RenderParams(w, h, 5, spp, L, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0)
    gpu_ctx.render_device(cam, P, tex1=tex1)
    torch.cuda.synchronize()
    return cam, tex1, P


# bound on device bytes / oracle bytes per ray. 1.10 is the bar (round-1 review). The Cornell box is the stated exception: 32 triangles in
# 3 nodes, every ray hits, and a node uncovers up to 24 triangles at once — the device tests them one per step while its node tests go on
# with the hit distance of a few steps ago, so it tests 3.9 triangles per ray where the oracle's test-at-once traversal tests 2.5 (node
# tests are equal). The policy is the measured optimum in TIME on every scene incl. this one (profiles/r02h_drain_first.jsonl).
@pytest.mark.parametrize("scene,kw,bound", [("terrain", TERRAIN, 1.10), ("cornell", {}, 1.25),
                                            ("instanced_rocks", dict(n_instances=60, rock_res=17, env_size=(64, 32)), 1.10)])
def test_device_work_is_within_ten_percent_of_the_oracle_count(gpu_ctx, scene, kw, bound):
    t = tables(scene, **kw)
    upload_scene(gpu_ctx, t)
    o = make_oracle(t)
    acc = HostAccel(t)
    tot = {"rays": 0, "nodes": 0, "tris": 0}
    for L in (2, 3, 4, 5):
        cam, tex1, P = _frame(gpu_ctx, t, L)
        rays, leaves = gpu_ctx.debug_ray_queue(L - 1)
        assert len(rays) > 0
        # the read-back rays are what the device traced: same hits from the oracle's walker on the product BVH8 and from mpt_trace
        ref, ref_t, rec, cnt = o.bvh8_trace(acc.view, rays, skip_records=None if acc.view.two_level else leaves)
        for k in tot:
            tot[k] += cnt[k]
        if L == 2:
            got = gpu_ctx.trace(rays)
            full, _, _, _ = o.bvh8_trace(acc.view, rays)  # mpt_trace knows nothing about the triangle a ray starts on
            same = (full["primitive_id"] == got["primitive_id"]) & (full["instance_id"] == got["instance_id"]) & (full["geometry_index"] == got["geometry_index"])
            assert 1.0 - same.mean() < 1e-4
            assert (full["barycentrics"].view(np.uint32)[same] == got["barycentrics"].view(np.uint32)[same]).all()
    gpu_ctx.set_profiling(1)  # counts every traversal kernel: the primary hits are computed beforehand
    gpu_ctx.reset_stats()
    _frame(gpu_ctx, t, 5, tex1=tex1)
    cs = gpu_ctx.stats()
    gpu_ctx.set_profiling(0)
    assert cs.rays_continuation == tot["rays"]  # bounce by bounce the same rays (the RNG is keyed on pixel, frame, sample)
    dn, dt = cs.nodes_visited / cs.rays_continuation, cs.tris_tested / cs.rays_continuation
    on, ot = tot["nodes"] / tot["rays"], tot["tris"] / tot["rays"]
    print(f"{scene}: nodes/ray device {dn:.3f} oracle {on:.3f}; triangles/ray device {dt:.3f} oracle {ot:.3f}")
    assert dn <= 1.10 * on
    assert 80 * dn + 48 * dt <= bound * (80 * on + 48 * ot)  # the algorithmic bytes per ray bench.py reports
    acc.close()


def test_ray_queue_readback_matches_the_counters(gpu_ctx):
    t = tables("terrain", **TERRAIN)
    upload_scene(gpu_ctx, t)
    gpu_ctx.reset_stats()
    _frame(gpu_ctx, t, 2)
    rays, leaves = gpu_ctx.debug_ray_queue(1)
    st = gpu_ctx.stats()
    assert len(rays) == st.rays_continuation > 0
    assert np.isfinite(rays["origin"]).all() and np.allclose(np.linalg.norm(rays["direction"], axis=1), 1.0, atol=1e-4)
    assert (rays["tmin"] == 0).all() and (rays["tmax"] == np.float32(1e7)).all()
    assert (leaves < t.triangle_count).all()  # every continuation ray of bounce 1 leaves a triangle record of the scene
