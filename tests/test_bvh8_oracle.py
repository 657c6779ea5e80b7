"""The oracle's BVH8 walker (oracle/oracle_bvh8.hpp) over the PRODUCT's acceleration structure (mpt_host_accel_build) against the
oracle's independent BVH2 traversal: same hits bit for bit, and sane work counters. CPU only.

This pins three things without a GPU: (1) the host builder emits a tree that contains every triangle and is walkable in the octant
order its layout encodes, (2) the instance records / TLAS of two-level scenes are consistent with the scene description, (3) the
counters bench.py uses as the roofline numerator come from a traversal that finds the right hits."""
import numpy as np
import pytest

from merian_b200 import abi, synthetic
from merian_b200.host_accel import HostAccel

from helpers import make_oracle, tables


def _camera_rays(o, cam, w, h):
    import ctypes as C
    rays = np.zeros(w * h, abi.RAY_DTYPE)
    for y in range(h):
        for x in range(w):
            r = abi.Ray()
            o.l.oracle_camera_ray(C.byref(cam), C.c_uint32(x), C.c_uint32(y), C.c_uint32(w), C.c_uint32(h), C.byref(r))
            i = y * w + x
            rays["origin"][i] = r.origin[:]
            rays["tmin"][i] = r.tmin
            rays["direction"][i] = r.direction[:]
            rays["tmax"][i] = r.tmax
    return rays


def _assert_same_hits(ref, got):
    for f in ("instance_id", "instance_index", "primitive_id", "geometry_index"):
        assert (ref[f] == got[f]).all(), f
    assert (ref["barycentrics"].view(np.uint32) == got["barycentrics"].view(np.uint32)).all()


@pytest.mark.parametrize("scene,kw", [("cornell", {}), ("terrain", dict(nx=120, ny=60, env_size=(64, 32))),
                                      ("instanced_rocks", dict(n_instances=40, rock_res=13, env_size=(64, 32)))])
def test_bvh8_walker_matches_bvh2_oracle(scene, kw):
    t = tables(scene, **kw)
    o = make_oracle(t)
    acc = HostAccel(t)
    cam = t.desc.camera.to_abi(0)
    w, h = 96, 54
    rays = np.concatenate([_camera_rays(o, cam, w, h), synthetic.incoherent_rays(cam, w, h)])
    ref, ref_t = o.trace(rays, return_t=True)
    got, got_t, rec, cnt = o.bvh8_trace(acc.view, rays)
    _assert_same_hits(ref, got)
    assert (ref_t.view(np.uint32) == got_t.view(np.uint32)).all()
    hit = got["instance_id"] != 0xFFFFFFFF
    assert hit.any()
    # the record index of a hit leads back to the same primitive
    tri = acc.tris[rec[hit]]
    assert (tri[:, 0, 3].view(np.uint32) == got["primitive_id"][hit]).all()
    assert (tri[:, 1, 3].view(np.uint32) == got["geometry_index"][hit]).all()
    assert cnt["rays"] == len(rays) and cnt["nodes"] >= cnt["rays"] and cnt["tris"] >= int(hit.sum())
    assert cnt["max_stack"] <= 64
    if acc.view.two_level:
        assert cnt["instance_entries"] > 0
    acc.close()


def test_bvh8_walker_skip_record_and_mask():
    t = tables("terrain", nx=120, ny=60, env_size=(64, 32))
    o = make_oracle(t)
    acc = HostAccel(t)
    cam = t.desc.camera.to_abi(0)
    rays = _camera_rays(o, cam, 64, 36)
    got, got_t, rec, cnt = o.bvh8_trace(acc.view, rays)
    hit = got["instance_id"] != 0xFFFFFFFF
    # skipping the record a ray hits makes it find something else (or nothing), never the same record
    got2, _, rec2, cnt2 = o.bvh8_trace(acc.view, rays, skip_records=rec)
    assert (rec2[hit] != rec[hit]).all()
    assert cnt2["tris"] < cnt["tris"] + int(hit.sum())
    # instance mask 0: nothing is hit, nothing is visited
    got3, _, _, cnt3 = o.bvh8_trace(acc.view, rays, mask=0)
    assert (got3["instance_id"] == 0xFFFFFFFF).all() and cnt3["nodes"] == 0
    acc.close()


def test_host_accel_covers_every_primitive():
    t = tables("terrain", nx=60, ny=30, env_size=(64, 32))
    acc = HostAccel(t)
    p2t = acc.prim_to_tri
    assert len(p2t) == acc.view.triangle_count == t.triangle_count
    assert np.array_equal(np.sort(p2t), np.arange(len(p2t), dtype=np.uint32))  # a permutation: every primitive has exactly one record
    nodes = acc.nodes
    # every leaf child's records lie inside the record array and leaf ranges are disjoint
    covered = np.zeros(acc.view.triangle_count, np.int32)
    for nd in nodes:
        lb, first = int(nd["leaf_bits"]), int(nd["prim_base"])
        assert lb >> 24 == 0
        for s in range(8):
            m = (lb >> (3 * s)) & 7
            if m == 0:
                continue
            assert not (int(nd["imask"]) >> s) & 1  # a slot is an inner child or a leaf, never both
            count = {1: 1, 3: 2, 7: 3}[m]
            covered[first:first + count] += 1  # records of the leaf slots lie back to back in slot order
            first += count
    assert (covered == 1).all()
    acc.close()
