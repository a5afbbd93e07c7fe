"""Known-answer tests for the oracle's TSDF integration restatement."""
import numpy as np

from voxgraph_b200 import synth

VS = 0.2
IDENT = np.array([1, 0, 0, 0, 0, 0, 0], np.float32)


def _voxel(layer_export, vps, g):
    idx, d, w = layer_export
    b = tuple(np.asarray(g) >> 4); l = np.asarray(g) & (vps - 1)
    for k in range(idx.shape[0]):
        if tuple(idx[k]) == b:
            lin = l[0] + vps * (l[1] + vps * l[2])
            return d[k, lin], w[k, lin]
    return None


def test_ka8_single_ray(oracle):
    """One ray (nearly) along +x from the origin to (d,0,0)."""
    layer = oracle.Layer(VS, 16)
    cfg = oracle.tsdf_config(use_sparsity_compensation_factor=0)
    d = 5.03
    # a hair off-axis so no DDA axis is degenerate
    p = np.array([[d, 1e-4, 2e-4]], np.float32)
    # the ray starts inside voxel (0,0,0): shift the sensor to the voxel centre
    T = IDENT.copy(); T[4:] = [0.1, 0.1, 0.1]
    st = oracle.tsdf_integrate(layer, cfg, T, p)
    assert st.rays_valid == 1 and st.rays_cast == 1
    trunc = 0.6
    end_x = int(np.floor((0.1 + d + trunc) / VS + 1e-6))
    assert st.voxel_updates == end_x + 1
    ex = layer.export()
    for i in range(end_x + 1):
        got = _voxel(ex, 16, (i, 0, 0))
        assert got is not None
        sdf = (0.1 + d) - (i + 0.5) * VS
        w_exp = 1.0
        if sdf < -VS:
            w_exp = max(0.0, (trunc + sdf) / (trunc - VS))
        if w_exp < 1e-6:
            assert got[1] == 0.0
            continue
        assert abs(got[1] - w_exp) < 1e-4, (i, got, w_exp)
        assert abs(got[0] - np.clip(sdf, -trunc, trunc)) < 1e-4, (i, got, sdf)
    # integrating the same scan again: same distance, doubled weight
    oracle.tsdf_integrate(layer, cfg, T, p)
    ex2 = layer.export()
    for i in range(end_x + 1):
        a = _voxel(ex, 16, (i, 0, 0)); b = _voxel(ex2, 16, (i, 0, 0))
        assert abs(b[1] - 2 * a[1]) < 1e-5 and abs(b[0] - a[0]) < 1e-5


def test_ka9_dda_visits_intersected_voxels(oracle):
    rs = np.random.RandomState(0)
    inv = float(np.float32(1.0 / np.float64(np.float32(VS))))
    for _ in range(200):
        o = rs.uniform(-3, 3, 3).astype(np.float32)
        p = (o + rs.uniform(-6, 6, 3)).astype(np.float32)
        idx, n = oracle.raycast(o, p, inv, 0.6)
        assert n == len(idx)
        start = np.floor(o * np.float32(inv) + np.float32(1e-6)).astype(np.int64)
        assert np.array_equal(idx[0], start)
        # consecutive indices differ by exactly one step along one axis
        dif = np.abs(np.diff(idx, axis=0)).sum(1)
        assert np.all(dif == 1)
        # count = Manhattan index distance + 1
        u = (p - o) / np.linalg.norm(p - o)
        e = (p + u * np.float32(0.6)).astype(np.float32)
        end = np.floor(e * np.float32(inv) + np.float32(1e-6)).astype(np.int64)
        assert n == np.abs(end - start).sum() + 1
        assert np.array_equal(idx[-1], end)
        # every visited voxel's cube is intersected by the segment (within tolerance)
        seg_o = o.astype(np.float64) * inv; seg_d = (e.astype(np.float64) - o) * inv
        for g in idx[:: max(1, len(idx) // 8)]:
            lo = g - 1e-3; hi = g + 1 + 1e-3
            with np.errstate(divide="ignore"):
                t1 = (lo - seg_o) / seg_d; t2 = (hi - seg_o) / seg_d
            tmin = np.minimum(t1, t2).max(); tmax = np.maximum(t1, t2).min()
            assert tmax >= max(tmin, 0) - 1e-6 and tmin <= 1 + 1e-6


def test_clearing_and_invalid_rays(oracle):
    layer = oracle.Layer(VS, 16)
    cfg = oracle.tsdf_config()
    pts = np.array([[0.05, 0.0, 0.0],     # < min_ray_length -> skipped
                    [30.0, 0.3, 0.2],     # > max_ray_length -> clearing ray, 16 m long
                    [4.0, 0.1, 0.1]], np.float32)
    st = oracle.tsdf_integrate(layer, cfg, IDENT, pts)
    assert st.rays_valid == 2
    idx, d, w = layer.export()
    # the clearing ray never writes a negative distance
    bx = idx[:, 0]
    far = bx >= 2  # x >= 6.4 m: only the clearing ray reaches
    assert far.any() and (d[far][w[far] > 0] > 0).all()
    x_end = 16.0 * 30.0 / np.linalg.norm([30.0, 0.3, 0.2])
    assert idx[:, 0].max() == int(np.floor(x_end / (16 * VS)))
    cfg2 = oracle.tsdf_config(allow_clear=0)
    layer2 = oracle.Layer(VS, 16)
    st2 = oracle.tsdf_integrate(layer2, cfg2, IDENT, pts)
    assert st2.rays_valid == 1


def test_fast_mode_is_subset_of_simple(oracle):
    world = synth.make_world(3, size_xy=(30.0, 30.0), n_clutter=40, n_walls=4)
    pose = np.array([15.0, 15.0, 1.2, 0.4])
    pts = synth.lidar_scan(world, pose, n_beams=16, n_azimuth=256, seed=1)
    assert pts.shape[0] > 2000
    T = synth.pose_to_T([0, 0, 0, 0])
    simple = oracle.Layer(VS, 16); fast = oracle.Layer(VS, 16)
    s1 = oracle.tsdf_integrate(simple, oracle.tsdf_config(mode=0), T, pts)
    s2 = oracle.tsdf_integrate(fast, oracle.tsdf_config(mode=1), T, pts)
    assert s2.rays_cast <= s1.rays_cast and s2.voxel_updates < s1.voxel_updates
    i1, d1, w1 = simple.export(); i2, d2, w2 = fast.export()
    set1 = set(map(tuple, i1))
    assert all(tuple(b) in set1 for b in i2)
    # surface voxels (|d| < trunc, observed) of fast agree in sign with simple
    m1 = {tuple(b): k for k, b in enumerate(i1)}
    agree = tot = 0
    for k, b in enumerate(i2):
        j = m1[tuple(b)]
        m = (w2[k] > 0) & (np.abs(d2[k]) < 0.3) & (w1[j] > 0)
        agree += (np.sign(d2[k][m]) == np.sign(d1[j][m])).sum(); tot += m.sum()
    assert tot > 100 and agree / tot > 0.9
