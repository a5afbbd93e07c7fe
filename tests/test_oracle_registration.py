"""Known-answer tests pinning the CPU oracle's registration path."""
import json
import os

import numpy as np
import pytest

from voxgraph_b200 import synth

VS, VPS = 0.2, 16


def _plane_layer(oracle, n, c, rng=((-2, 2), (-2, 2), (-1, 1))):
    idx, d, w = synth.plane_layer_blocks(n, c, VS, VPS, rng)
    return oracle.Layer.from_blocks(VS, VPS, idx, d, w)


def _rand_points(rs, k, lo=-3.0, hi=3.0):
    xyz = rs.uniform(lo, hi, (k, 3)).astype(np.float32)
    xyz[:, 2] = rs.uniform(-1.0, 1.0, k)
    return xyz


def _analytic(xyz, dist, w, ref, read, n, c):
    """Closed forms of registration_cost_function.cpp:128-239 for a planar reading field."""
    xyz = xyz.astype(np.float64)
    Rr, Re = synth.rot_z(ref[3]), synth.rot_z(read[3])
    pm = xyz @ Rr.T + ref[:3]
    pe = (pm - read[:3]) @ Re
    d = pe @ n - c
    r = (dist - d) * w
    ce, se = np.cos(read[3]), np.sin(read[3])
    cemo, semo = np.cos(read[3] - ref[3]), np.sin(read[3] - ref[3])
    xi, yi = xyz[:, 0], xyz[:, 1]
    k = len(xi)
    Aref = np.zeros((k, 3, 4)); Aread = np.zeros((k, 3, 4))
    Aref[:, 0, 0] = ce; Aref[:, 0, 1] = se; Aref[:, 0, 3] = xi * semo - yi * cemo
    Aref[:, 1, 0] = -se; Aref[:, 1, 1] = ce; Aref[:, 1, 3] = xi * cemo + yi * semo
    Aref[:, 2, 2] = 1
    dxe, dye = read[0] - ref[0], read[1] - ref[1]
    Aread[:, 0, 0] = -ce; Aread[:, 0, 1] = -se
    Aread[:, 0, 3] = -xi * semo + yi * cemo + dxe * se - dye * ce
    Aread[:, 1, 0] = se; Aread[:, 1, 1] = -ce
    Aread[:, 1, 3] = -xi * cemo - yi * semo + dxe * ce + dye * se
    Aread[:, 2, 2] = -1
    Jr = -w[:, None] * np.einsum("j,kjc->kc", n, Aref)
    Je = -w[:, None] * np.einsum("j,kjc->kc", n, Aread)
    f = k / w.sum()
    return r * f, Jr * f, Je * f


def test_ka1_plane_closed_form(oracle):
    rs = np.random.RandomState(0)
    n = np.array([0.3, -0.5, 0.81]); n /= np.linalg.norm(n)
    c = 0.17
    layer = _plane_layer(oracle, n, c)
    xyz = _rand_points(rs, 500)
    dist = rs.uniform(-0.2, 0.2, 500).astype(np.float32)
    w = rs.uniform(0.5, 3.0, 500).astype(np.float32)
    ref = np.array([0.4, -0.2, 0.1, 0.3]); read = np.array([-0.3, 0.5, -0.05, -0.4])
    ok, r, jr, je = oracle.reg_evaluate(layer, xyz, dist, w, ref, read)
    assert ok
    # all points must have interpolated (layer covers +-6.4 m)
    ra, jra, jea = _analytic(xyz, dist.astype(np.float64), w.astype(np.float64), ref, read, n, c)
    np.testing.assert_allclose(r, ra, rtol=0, atol=2e-5)
    np.testing.assert_allclose(jr, jra, rtol=0, atol=5e-5)
    np.testing.assert_allclose(je, jea, rtol=0, atol=5e-5)


@pytest.mark.parametrize("mono", ["x", "y", "z", "xy", "yz", "zx", "xyz"])
def test_ka2_trilinear_monomials(oracle, mono):
    """Trilinear interpolation reproduces each monomial exactly -> pins corner order vs B1
    (registration_cost_function.h:73-81) and the q-vector order."""
    fns = {"x": lambda p: p[:, 0], "y": lambda p: p[:, 1], "z": lambda p: p[:, 2],
           "xy": lambda p: p[:, 0] * p[:, 1], "yz": lambda p: p[:, 1] * p[:, 2],
           "zx": lambda p: p[:, 2] * p[:, 0], "xyz": lambda p: p[:, 0] * p[:, 1] * p[:, 2]}
    grads = {"x": lambda p: np.stack([np.ones(len(p)), 0 * p[:, 0], 0 * p[:, 0]], -1),
             "y": lambda p: np.stack([0 * p[:, 0], np.ones(len(p)), 0 * p[:, 0]], -1),
             "z": lambda p: np.stack([0 * p[:, 0], 0 * p[:, 0], np.ones(len(p))], -1),
             "xy": lambda p: np.stack([p[:, 1], p[:, 0], 0 * p[:, 0]], -1),
             "yz": lambda p: np.stack([0 * p[:, 0], p[:, 2], p[:, 1]], -1),
             "zx": lambda p: np.stack([p[:, 2], 0 * p[:, 0], p[:, 0]], -1),
             "xyz": lambda p: np.stack([p[:, 1] * p[:, 2], p[:, 0] * p[:, 2], p[:, 0] * p[:, 1]], -1)}
    idx, d, w = synth.field_layer_blocks(fns[mono], VS, VPS, ((-1, 1), (-1, 1), (-1, 1)))
    layer = oracle.Layer.from_blocks(VS, VPS, idx, d, w)
    rs = np.random.RandomState(1)
    xyz = rs.uniform(-2.5, 2.5, (300, 3)).astype(np.float32)
    zero = np.zeros(4)
    ok, r, jr, je = oracle.reg_evaluate(layer, xyz, np.zeros(300, np.float32),
                                        np.ones(300, np.float32), zero, zero)
    assert ok
    p = xyz.astype(np.float64)
    np.testing.assert_allclose(-r, fns[mono](p), atol=3e-5)
    # J_ref[:, 0:3] = -grad (identity poses: cos_e = 1, sin_e = 0)
    np.testing.assert_allclose(-jr[:, :3], grads[mono](p), atol=2e-4)
    np.testing.assert_allclose(je[:, :3], grads[mono](p), atol=2e-4)


def test_ka3_finite_differences(oracle, pair_scene):
    """Central differences of the residual vector vs the analytic Jacobians — the check the
    reference's NumericDiff toggle performs (submap_registration_helper.cpp:50-57)."""
    s0, s1 = pair_scene.submaps
    layer = oracle.Layer.from_blocks(s1.voxel_size, s1.vps, s1.block_idx, s1.distance, s1.weight)
    ref = pair_scene.poses_init[0].copy(); read = pair_scene.poses_init[1].copy()
    ok, r, jr, je = oracle.reg_evaluate(layer, s0.points_xyz, s0.points_distance,
                                        s0.points_weight, ref, read)
    assert ok
    h = 2e-3
    valid = np.abs(jr).sum(1) > 0
    assert valid.sum() > 200
    good_total, n_total = 0, 0
    for blk, J in ((0, jr), (1, je)):
        for c in range(4):
            pp = [ref.copy(), read.copy()]; pm = [ref.copy(), read.copy()]
            pp[blk][c] += h; pm[blk][c] -= h
            _, rp, _, _ = oracle.reg_evaluate(layer, s0.points_xyz, s0.points_distance,
                                              s0.points_weight, pp[0], pp[1], jacobians=False)
            _, rm, _, _ = oracle.reg_evaluate(layer, s0.points_xyz, s0.points_distance,
                                              s0.points_weight, pm[0], pm[1], jacobians=False)
            num = (rp - rm) / (2 * h)
            # piecewise-trilinear field: FD is only meaningful where no cell boundary or
            # correspondence change is crossed; require agreement on the bulk
            err = np.abs(num[valid] - J[valid, c])
            scale = np.abs(J[valid, c]) + 1.0
            good_total += (err < 0.05 * scale).sum(); n_total += valid.sum()
    assert good_total / n_total > 0.9


def test_ka4_sympy_golden_pose_jacobians(oracle):
    """3x4 pose Jacobians vs golden vectors lambdified from the reference's own sympy
    derivation (voxgraph/scripts/jacobians_xyz_yaw.py, via tests/golden/make_jacobian_golden.py).
    A reading field d = n.p makes pInterp_pr = n, so rows of the 3x4 matrices are read off
    with n = e_x, e_y, e_z."""
    path = os.path.join(os.path.dirname(__file__), "golden", "pose_jacobians.json")
    cases = json.load(open(path))["cases"]
    assert len(cases) >= 16
    for case in cases:
        ref = np.array(case["ref"]); read = np.array(case["read"]); p = np.array(case["point"])
        Aref = np.array(case["dTp_dref"]); Aread = np.array(case["dTp_dread"])
        # the point must land inside the layer
        for row, n in enumerate(np.eye(3)):
            layer = _plane_layer(oracle, n, 0.0, ((-3, 3), (-3, 3), (-2, 2)))
            ok, r, jr, je = oracle.reg_evaluate(layer, p[None].astype(np.float32),
                                                np.zeros(1, np.float32), np.ones(1, np.float32),
                                                ref, read)
            assert ok
            np.testing.assert_allclose(-jr[0], Aref[row], atol=2e-5)
            np.testing.assert_allclose(-je[0], Aread[row], atol=2e-5)


def test_ka6_normalisation_and_zero_weight(oracle):
    n = np.array([0.0, 0.0, 1.0])
    layer = _plane_layer(oracle, n, 0.0)
    rs = np.random.RandomState(3)
    xyz = _rand_points(rs, 64)
    dist = np.zeros(64, np.float32)
    w = rs.uniform(1, 2, 64).astype(np.float32)
    pose = np.zeros(4)
    ok, r1, j1, _ = oracle.reg_evaluate(layer, xyz, dist, w, pose, pose)
    ok2, r2, j2, _ = oracle.reg_evaluate(layer, xyz, dist, 3 * w, pose, pose)
    assert ok and ok2
    # scaling all weights by s leaves r (= r_raw * s * K/(s*sum w)) unchanged
    np.testing.assert_allclose(r1, r2, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(j1, j2, rtol=1e-6, atol=1e-9)
    # raw residual z*w times K / sum(w)
    np.testing.assert_allclose(r1, -(xyz[:, 2].astype(np.float64) * w) * 64 / w.astype(np.float64).sum(),
                               atol=2e-5)
    ok3, _, _, _ = oracle.reg_evaluate(layer, xyz, dist, np.zeros(64, np.float32), pose, pose)
    assert not ok3  # Evaluate returns false (cpp:273)


def test_ka7_no_correspondence(oracle):
    n = np.array([0.0, 0.0, 1.0])
    idx, d, w = synth.plane_layer_blocks(n, 0.0, VS, VPS, ((0, 0), (0, 0), (0, 0)))
    w = w.copy()
    # voxel (5,5,5) unobserved
    w[0, 5 + 16 * (5 + 16 * 5)] = 0.0
    layer = oracle.Layer.from_blocks(VS, VPS, idx, d, w)
    pts = np.array([[2.03, 2.03, 2.03],    # fine
                    [10.0, 1.0, 1.0],      # block missing
                    [3.15, 1.0, 1.0],      # +x neighbour block missing (needs voxel 16)
                    [1.05, 1.05, 1.05],    # touches the unobserved voxel (5,5,5)
                    [0.05, 1.0, 1.0]],     # lower neighbour would be in block -1
                   np.float32)
    wt = np.array([1, 2, 3, 4, 5], np.float32)
    pose = np.zeros(4)
    ok, r, jr, je = oracle.reg_evaluate(layer, pts, np.zeros(5, np.float32), wt, pose, pose,
                                        no_correspondence_cost=0.25)
    assert ok
    f = 5 / wt.sum()
    np.testing.assert_allclose(r[1:], wt[1:].astype(np.float64) * 0.25 * (5.0 / 15.0), rtol=1e-12)
    assert np.all(jr[1:] == 0) and np.all(je[1:] == 0)
    assert abs(r[0] - (-2.03 * float(f))) < 1e-5 and np.any(jr[0] != 0)


def test_ka10_index_math_bit_exact(oracle):
    rs = np.random.RandomState(5)
    inv = float(np.float32(1.0 / np.float64(np.float32(0.2))))
    for _ in range(2000):
        p = (rs.uniform(-200, 200, 3)).astype(np.float32)
        got = oracle.grid_index_from_point(p, inv)
        exp = np.floor(p * np.float32(inv) + np.float32(1e-6)).astype(np.int32)
        assert np.array_equal(got, exp)
    for _ in range(2000):
        g = rs.randint(-100000, 100000, 3).astype(np.int64)
        b, l = oracle.block_and_local_from_global(g, 16)
        assert np.array_equal(b, g >> 4) and np.array_equal(l, g & 15)


def test_interp_neighbour_layout(oracle):
    """Corner i of getVoxelsAndQVector has x = bit2, y = bit1, z = bit0 (A.3)."""
    idx, d, w = synth.field_layer_blocks(lambda p: p[:, 0] * 100 + p[:, 1] * 10 + p[:, 2], VS, VPS,
                                         ((0, 1), (0, 1), (0, 1)))
    layer = oracle.Layer.from_blocks(VS, VPS, idx, d, w)
    res = layer.interp([3.15, 3.15, 3.15])  # crosses all three block faces
    assert res["ok"]
    assert np.array_equal(res["base_block"], [0, 0, 0]) and np.array_equal(res["base_voxel"], [15, 15, 15])
    d = res["distances"].astype(np.float64)
    np.testing.assert_allclose(d[4] - d[0], 100 * 0.2, rtol=1e-4)
    np.testing.assert_allclose(d[2] - d[0], 10 * 0.2, rtol=1e-4)
    np.testing.assert_allclose(d[1] - d[0], 0.2, rtol=1e-3)
    assert len(set(res["slots"].tolist())) == 8
    np.testing.assert_allclose(res["q"][1:4], [0.25, 0.25, 0.25], atol=1e-4)


def test_transform_matches_yaw_rotation(oracle):
    rs = np.random.RandomState(7)
    for _ in range(50):
        ref = rs.uniform(-5, 5, 4); read = rs.uniform(-5, 5, 4)
        ref[3] = rs.uniform(-3.1, 3.1); read[3] = rs.uniform(-3.1, 3.1)
        T, trig = oracle.reg_pose_setup(ref, read)
        p = rs.uniform(-10, 10, 3)
        got = oracle.T_transform(T, p)
        exp = (p @ synth.rot_z(ref[3]).T + ref[:3] - read[:3]) @ synth.rot_z(read[3])
        np.testing.assert_allclose(got, exp, atol=2e-5)
        assert T[1] == 0 and T[2] == 0
        np.testing.assert_allclose(trig[:4], [np.cos(read[3]), np.sin(read[3]),
                                              np.cos(read[3] - ref[3]), np.sin(read[3] - ref[3])],
                                   atol=1e-6)


def test_random_field_matches_bruteforce_trilinear(oracle):
    """Arbitrary (random) voxel data: the oracle's interpolated distance equals a brute-force
    numpy trilinear interpolation with explicit corner weights -> pins corner/q ordering on data
    that is not a low-order polynomial, including block-face crossings."""
    rs = np.random.RandomState(9)
    idx = np.array([[bx, by, bz] for bz in (-1, 0) for by in (-1, 0) for bx in (-1, 0)], np.int32)
    d = rs.normal(size=(8, VPS ** 3)).astype(np.float32)
    w = np.ones_like(d)
    layer = oracle.Layer.from_blocks(VS, VPS, idx, d, w)
    # dense lookup table of the same data
    dense = np.zeros((32, 32, 32), np.float32)
    for k, (bx, by, bz) in enumerate(idx):
        blk = d[k].reshape(VPS, VPS, VPS)          # [z][y][x]
        dense[(bx + 1) * 16:(bx + 2) * 16, (by + 1) * 16:(by + 2) * 16, (bz + 1) * 16:(bz + 2) * 16] = \
            blk.transpose(2, 1, 0)
    xyz = rs.uniform(-3.0, 3.0, (2000, 3)).astype(np.float32)
    ok, r, jr, je = oracle.reg_evaluate(layer, xyz, np.zeros(2000, np.float32), np.ones(2000, np.float32),
                                        np.zeros(4), np.zeros(4))
    assert ok
    p = xyz.astype(np.float64) / 0.2 - 0.5 + 16      # continuous voxel-centre coordinates in `dense`
    i0 = np.floor(p).astype(int); f = p - i0
    exp = np.zeros(2000); grad = np.zeros((2000, 3))
    inside = np.all((i0 >= 0) & (i0 < 31), axis=1)
    for cx in (0, 1):
        for cy in (0, 1):
            for cz in (0, 1):
                wx = np.where(cx, f[:, 0], 1 - f[:, 0]); wy = np.where(cy, f[:, 1], 1 - f[:, 1])
                wz = np.where(cz, f[:, 2], 1 - f[:, 2])
                ii = np.clip(i0 + [cx, cy, cz], 0, 31)
                val = dense[ii[:, 0], ii[:, 1], ii[:, 2]].astype(np.float64)
                exp += wx * wy * wz * val
                grad[:, 0] += (1 if cx else -1) * wy * wz * val / 0.2
                grad[:, 1] += wx * (1 if cy else -1) * wz * val / 0.2
                grad[:, 2] += wx * wy * (1 if cz else -1) * val / 0.2
    has = np.abs(jr).sum(1) > 0
    # points whose 8 corners lie inside the 2x2x2 blocks interpolate; the others have no correspondence
    assert has[inside].mean() > 0.99 and inside.sum() > 1500
    m = inside & has
    np.testing.assert_allclose(-r[m], exp[m], atol=2e-5)
    np.testing.assert_allclose(-jr[m, :3], grad[m], atol=5e-4)
