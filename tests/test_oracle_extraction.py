"""Known-answer tests of the oracle's findIsosurfaceVertices / Interpolator::getVoxel restatement
(voxgraph_submap.cpp:203-243; voxblox MeshIntegrator / MarchingCubes / createConnectedMesh are
restated from the published algorithm: parity unpinned against the binaries)."""
import numpy as np

from voxgraph_b200 import synth

VS, VPS = 0.2, 16


def _layer(oracle, idx, d, w):
    return oracle.Layer.from_blocks(VS, VPS, idx, d, w)


def test_axis_aligned_plane_vertices_lie_on_the_plane(oracle):
    """KA: d = z - 0.37 is reproduced exactly by linear interpolation: one vertex per voxel column,
    on the plane, distance 0, weight = the constant layer weight; border columns whose trilinear
    neighbourhood leaves the layer are dropped (31 x 31 of the 32 x 32 columns)."""
    idx, d, w = synth.plane_layer_blocks([0, 0, 1.0], 0.37, VS, VPS, ((0, 1), (0, 1), (-1, 0)))
    L = _layer(oracle, idx, d, 2.0 * w)
    xyz, dd, ww, blk = oracle.find_isosurface_vertices(L, 1.0)
    assert len(xyz) == 31 * 31
    assert np.abs(xyz[:, 2] - 0.37).max() < 1e-6
    assert np.abs(dd).max() < 1e-6 and np.allclose(ww, 2.0)
    # vertices sit above voxel centres (x, y) = (i + 0.5) * voxel_size
    fx = xyz[:, 0] / VS - 0.5
    assert np.abs(fx - np.round(fx)).max() < 1e-4
    assert sorted(map(tuple, blk)) == [(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 1, 0)]


def test_min_weight_gates_cubes(oracle):
    idx, d, w = synth.plane_layer_blocks([0, 0, 1.0], 0.37, VS, VPS, ((0, 1), (0, 1), (-1, 0)))
    L = _layer(oracle, idx, d, w)          # weight 1 is NOT > min_weight 1 (getSdfIfValid)
    xyz, _, _, _ = oracle.find_isosurface_vertices(L, 1.0)
    assert len(xyz) == 0
    xyz, _, _, _ = oracle.find_isosurface_vertices(L, 0.5)
    assert len(xyz) == 31 * 31


def test_sphere_vertices_and_merge_rule(oracle):
    """KA: sphere of radius 1.3 m: every vertex within the linear-interpolation error of the surface,
    no two surviving vertices share a round(v / (0.5 voxel)) bucket, interpolated distance ~ 0."""
    c = np.array([1.7, 1.5, 1.6]); R = 1.3
    idx, d, w = synth.field_layer_blocks(lambda p: np.linalg.norm(p - c, axis=-1) - R, VS, VPS,
                                         ((0, 1), (0, 1), (0, 1)), weight=3.0)
    d = np.clip(d, -0.6, 0.6).astype(np.float32)
    L = _layer(oracle, idx, d, w)
    xyz, dd, ww, blk = oracle.find_isosurface_vertices(L, 1.0)
    assert len(xyz) > 500
    r = np.linalg.norm(xyz.astype(np.float64) - c, axis=1)
    assert np.abs(r - R).max() < 0.02          # chord error of linear interpolation at 0.2 m voxels
    assert np.abs(dd).max() < 0.02 and np.allclose(ww, 3.0, atol=1e-5)
    thr = np.float32(0.5 * np.float64(np.float32(VS)))
    keys = np.round(xyz.astype(np.float64) * (1.0 / np.float64(thr))).astype(np.int64)
    assert len(np.unique(keys, axis=0)) == len(keys)
    # every crossing edge is represented: vertex count is within the 4-fold sharing bound
    assert len(xyz) < 4 * np.pi * R * R / (VS * VS) * 3


def test_interp_voxel_trilinear_exact_on_affine_field(oracle):
    idx, d, w = synth.field_layer_blocks(lambda p: 0.3 * p[..., 0] - 0.2 * p[..., 1] + 0.5 * p[..., 2] - 1.0,
                                         VS, VPS, ((0, 1), (0, 1), (0, 1)), weight=2.5)
    L = _layer(oracle, idx, d, w)
    rs = np.random.RandomState(1)
    for _ in range(200):
        p = rs.uniform(0.4, 5.6, 3).astype(np.float32)
        ok, dd, ww = oracle.interp_voxel(L, p)
        assert ok
        assert abs(dd - (0.3 * p[0] - 0.2 * p[1] + 0.5 * p[2] - 1.0)) < 2e-5
        assert abs(ww - 2.5) < 1e-5
    ok, _, _ = oracle.interp_voxel(L, np.array([-0.5, 1.0, 1.0], np.float32))
    assert not ok


def test_relevant_voxels_subset_and_obb(oracle):
    idx, d, w = synth.plane_layer_blocks([0, 0, 1.0], 0.37, VS, VPS, ((0, 1), (0, 1), (-1, 0)))
    L = _layer(oracle, idx, d, 2.0 * w)
    xyz, dd, ww = oracle.find_relevant_voxels(L, 1.0, 0.3)
    assert len(xyz) > 0 and np.abs(dd).max() < 0.3
    ok, mn, mx = oracle.surface_obb(L, 1.0, 0.3)
    assert ok
    assert np.allclose(mn, xyz.min(0) - 0.5 * VS, atol=1e-6) and np.allclose(mx, xyz.max(0) + 0.5 * VS, atol=1e-6)
