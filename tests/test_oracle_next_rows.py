"""Oracle restatements of SURVEY §8f rows (checked ahead of their GPU versions):
findRelevantVoxelIndices, surface OBB / mission-frame AABB, overlapsWith."""
import numpy as np

from voxgraph_b200 import synth


def _layer(oracle, s):
    return oracle.Layer.from_blocks(s.voxel_size, s.vps, s.block_idx, s.distance, s.weight)


def test_find_relevant_voxels_matches_bruteforce(oracle, pair_scene):
    s = pair_scene.submaps[0]
    layer = _layer(oracle, s)
    xyz, d, w = oracle.find_relevant_voxels(layer, 1.0, 0.3)
    m = (s.weight.astype(np.float64) > 1.0) & (np.abs(s.distance.astype(np.float64)) < 0.3)
    assert len(d) == m.sum() > 1000
    bi, li = np.nonzero(m)                      # block-major, linear-index order == reference order
    np.testing.assert_array_equal(d, s.distance[bi, li])
    np.testing.assert_array_equal(w, s.weight[bi, li])
    vps = s.vps
    v = np.stack([li % vps, (li // vps) % vps, li // (vps * vps)], -1).astype(np.float32)
    bs = np.float32(s.voxel_size) * np.float32(vps)
    exp = s.block_idx[bi].astype(np.float32) * bs + (v + np.float32(0.5)) * np.float32(s.voxel_size)
    np.testing.assert_array_equal(xyz, exp.astype(np.float32))
    # every relevant voxel interpolates to (almost) its own distance when registered against itself
    ok, r, _, _ = oracle.reg_evaluate(layer, xyz, d, w, s.pose_gt, s.pose_gt, jacobians=False)
    assert ok and np.abs(r).max() < 1e-4


def test_surface_obb_and_aabb(oracle, pair_scene):
    s = pair_scene.submaps[1]
    layer = _layer(oracle, s)
    ok, mn, mx = oracle.surface_obb(layer, 1.0, 0.3)
    xyz, d, w = oracle.find_relevant_voxels(layer, 1.0, 0.3)
    assert ok
    half = np.float32(0.5) * np.float32(s.voxel_size)
    np.testing.assert_array_equal(mn, (xyz - half).min(0))
    np.testing.assert_array_equal(mx, (xyz + half).max(0))
    T = synth.pose_to_T(s.pose_gt)
    amin, amax = oracle.aabb_from_obb_and_pose(mn, mx, T)
    corners = np.array([[x, y, z] for x in (mn[0], mx[0]) for y in (mn[1], mx[1]) for z in (mn[2], mx[2])])
    world = corners @ synth.rot_z(s.pose_gt[3]).T + s.pose_gt[:3]
    np.testing.assert_allclose(amin, world.min(0), atol=1e-4)
    np.testing.assert_allclose(amax, world.max(0), atol=1e-4)
    ok2, _, _ = oracle.surface_obb(oracle.Layer(0.2, 16), 1.0, 0.3)
    assert not ok2


def test_overlap_matches_host_restatement(oracle):
    """The C restatement of overlapsWith agrees with the independent numpy restatement used to build
    the synthetic pair lists (voxgraph_b200/synth.py::submaps_overlap) on every pair of the scene."""
    sc = synth.make_scene(seed=4, n_submaps=8, n_points=500, radius=6.0, size_xy=(96.0, 64.0),
                          n_clutter=100, n_walls=8)
    layers = [_layer(oracle, s) for s in sc.submaps]
    n = len(sc.submaps)
    got, exp = [], []
    for i in range(n):
        for j in range(i + 1, n):
            a, b = sc.submaps[i], sc.submaps[j]
            pa, pb = sc.poses_init[i], sc.poses_init[j]
            bs = np.float32(a.voxel_size * a.vps)
            def aabb(s, p):
                lo = s.surface_blocks.min(0).astype(np.float32) * bs
                hi = (s.surface_blocks.max(0) + 1).astype(np.float32) * bs
                return oracle.aabb_from_obb_and_pose(lo, hi, synth.pose_to_T(p))
            got.append(oracle.submaps_overlap(aabb(a, pa), aabb(b, pb), synth.pose_to_T(pa), synth.pose_to_T(pb),
                                              a.surface_blocks, float(bs), layers[j]))
            exp.append(synth.submaps_overlap(a, b, pa, pb))
    assert got == exp and any(got) and not all(got)
