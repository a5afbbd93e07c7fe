"""Known-answer tests for the oracle's pose-graph path."""
import numpy as np
import pytest

from voxgraph_b200 import synth


def test_ka11_relative_pose_residual(oracle):
    L = np.diag([1.0, 1.0, 50.0, 50.0])
    A = np.array([1.0, 2.0, 0.5, 3.0]); B = np.array([2.0, 1.0, 0.7, -3.0])
    t_obs, yaw_obs = synth.relative_pose(A, B)
    r, ja, jb = oracle.relpose_evaluate(A, B, t_obs, yaw_obs, L)
    np.testing.assert_allclose(r, 0, atol=1e-12)
    # yaw wrap across +-pi: B.yaw - A.yaw = -6 -> wrapped 0.283
    r2, _, _ = oracle.relpose_evaluate(A, B, t_obs, 0.0, L)
    assert abs(r2[3] - 50.0 * (2 * np.pi - 6.0)) < 1e-9
    # finite-difference check of the analytic Jacobians
    rs = np.random.RandomState(0)
    Lf = np.linalg.cholesky(np.array([[2, .3, 0, 0], [.3, 1, .1, 0], [0, .1, 30, 1], [0, 0, 1, 20.]]))
    for _ in range(10):
        A = rs.uniform(-2, 2, 4); B = rs.uniform(-2, 2, 4)
        t = rs.uniform(-1, 1, 3); y = rs.uniform(-0.5, 0.5)
        r0, ja, jb = oracle.relpose_evaluate(A, B, t, y, Lf)
        for blk, J in ((0, ja), (1, jb)):
            for c in range(4):
                h = 1e-6
                P = [A.copy(), B.copy()]; M = [A.copy(), B.copy()]
                P[blk][c] += h; M[blk][c] -= h
                rp, _, _ = oracle.relpose_evaluate(P[0], P[1], t, y, Lf)
                rm, _, _ = oracle.relpose_evaluate(M[0], M[1], t, y, Lf)
                np.testing.assert_allclose((rp - rm) / (2 * h), J[:, c], atol=1e-6)


def test_sqrt_information_is_llt_lower(oracle):
    info = np.array([[4, 1, 0, 0], [1, 3, .5, 0], [0, .5, 2500, 10], [0, 0, 10, 2500.]])
    L = oracle.sqrt_information(info)
    np.testing.assert_allclose(L, np.linalg.cholesky(info), atol=1e-12)
    with pytest.raises(ValueError):
        oracle.sqrt_information(-np.eye(4))
    assert abs(oracle.normalize_angle(np.pi) - (-np.pi)) < 1e-15
    assert abs(oracle.normalize_angle(-np.pi) - (-np.pi)) < 1e-15
    assert abs(oracle.normalize_angle(3 * np.pi + 0.1) - (-np.pi + 0.1)) < 1e-12


def test_ka12_lm_consistent_loop(oracle):
    """Relative-pose edges forming a consistent loop: the solution is the exact chain, cost -> 0."""
    rs = np.random.RandomState(1)
    n = 8
    gt = np.stack([3 * np.cos(np.linspace(0, 2 * np.pi, n, endpoint=False)),
                   3 * np.sin(np.linspace(0, 2 * np.pi, n, endpoint=False)),
                   np.linspace(0, 0.5, n), np.linspace(-2.5, 2.5, n)], -1)
    g = oracle.Graph()
    L = np.diag([1.0, 1.0, 50.0, 50.0])
    for i in range(n):
        init = gt[i] + (0 if i == 0 else rs.normal(0, 0.3, 4))
        g.add_node(i, init, constant=(i == 0))
    for i in range(n):
        j = (i + 1) % n
        t, y = synth.relative_pose(gt[i], gt[j])
        g.add_relative(i, j, t, y, L)
    rc, s = g.solve(oracle.solver_options(parameter_tolerance=1e-12, function_tolerance=1e-16,
                                          max_num_iterations=100))
    assert rc == 0
    assert s.final_cost < 1e-16 * max(1.0, s.initial_cost) + 1e-18
    x = g.poses()
    dyaw = np.array([oracle.normalize_angle(a) for a in (x[:, 3] - gt[:, 3])])
    np.testing.assert_allclose(x[:, :3], gt[:, :3], atol=1e-7)
    np.testing.assert_allclose(dyaw, 0, atol=1e-7)
    # gradient/H consistency: H == J^T J is symmetric PSD
    ok, cost, grad, H = g.eval()
    assert ok and np.allclose(H, H.T) and np.linalg.eigvalsh(H).min() > -1e-9


def _layer_of(oracle, s):
    return oracle.Layer.from_blocks(s.voxel_size, s.vps, s.block_idx, s.distance, s.weight)


@pytest.mark.parametrize("pert", [(0.3, 0.0, 0.0, 0.0), (0.0, -0.6, 0.0, 0.0), (0.0, 0.0, 0.3, 0.0),
                                  (0.0, 0.0, 0.0, 0.1), (0.3, -0.6, 0.3, -0.2)])
def test_ka5_self_registration(oracle, pair_scene, pert):
    """registration_test_bench protocol (registration_test_bench.cpp:298-347, yaml:9-13): a submap
    registered against a copy of itself under a known perturbation converges back; reference pose
    constant, reading free (submap_registration_helper.cpp:29-70)."""
    s = pair_scene.submaps[0]
    layer = _layer_of(oracle, s)
    base = s.pose_gt
    g = oracle.Graph()
    g.add_node(0, base, constant=True)
    g.add_node(1, base + np.array(pert), constant=False)
    g.add_registration(0, 1, layer, s.points_xyz, s.points_distance, s.points_weight)
    ok, c0, _, _ = g.eval()
    assert ok and c0 > 0
    rc, summ = g.solve(oracle.solver_options(parameter_tolerance=3e-9, max_num_iterations=40))
    assert rc == 0
    err = g.poses()[1] - base
    assert np.abs(err[:3]).max() < 0.02 and abs(err[3]) < 0.005, (err, summ.final_cost)
    assert summ.final_cost < 1e-3 * c0


def test_ka5_zero_perturbation_cost(oracle, pair_scene):
    s = pair_scene.submaps[0]
    layer = _layer_of(oracle, s)
    ok, r, jr, je = oracle.reg_evaluate(layer, s.points_xyz, s.points_distance, s.points_weight,
                                        s.pose_gt, s.pose_gt)
    assert ok
    # points are the layer's own isosurface: interpolated distance == stored vertex distance
    assert np.abs(r).max() < 1e-4


def test_graph_eval_matches_blockwise(oracle, small_scene):
    """H and g assembled by vgo_graph_eval equal the sum of per-block J^T J / J^T r."""
    sc = small_scene
    g = oracle.Graph()
    layers = [_layer_of(oracle, s) for s in sc.submaps]
    for i, s in enumerate(sc.submaps):
        g.add_node(i, sc.poses_init[i], constant=(i == 0))
    L = oracle.sqrt_information(sc.odom_information)
    for (i, j, t, y) in sc.odometry:
        g.add_relative(i, j, t, y, L)
    for (i, j) in sc.pairs:
        a, b = sc.submaps[i], sc.submaps[j]
        g.add_registration(i, j, layers[j], a.points_xyz, a.points_distance, a.points_weight)
        g.add_registration(j, i, layers[i], b.points_xyz, b.points_distance, b.points_weight)
    ok, cost, grad, H = g.eval(num_threads=4)
    assert ok
    n = 4 * len(sc.submaps)
    Hx = np.zeros((n, n)); gx = np.zeros(n); cx = 0.0
    x = sc.poses_init
    for (i, j, t, y) in sc.odometry:
        r, ja, jb = oracle.relpose_evaluate(x[i], x[j], t, y, L)
        J = np.zeros((4, n)); J[:, 4 * i:4 * i + 4] = ja; J[:, 4 * j:4 * j + 4] = jb
        Hx += J.T @ J; gx += J.T @ r; cx += 0.5 * r @ r
    for (i, j) in sc.pairs:
        for (a, b) in ((i, j), (j, i)):
            sa = sc.submaps[a]
            ok1, r, jr, je = oracle.reg_evaluate(layers[b], sa.points_xyz, sa.points_distance,
                                                 sa.points_weight, x[a], x[b])
            assert ok1
            J = np.zeros((len(r), n)); J[:, 4 * a:4 * a + 4] = jr; J[:, 4 * b:4 * b + 4] = je
            Hx += J.T @ J; gx += J.T @ r; cx += 0.5 * r @ r
    np.testing.assert_allclose(cost, cx, rtol=1e-10)
    np.testing.assert_allclose(grad, gx, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(H, Hx, rtol=1e-9, atol=1e-8)
    # threaded == serial
    ok, cost1, grad1, H1 = g.eval(num_threads=1)
    assert cost1 == cost and np.array_equal(grad1, grad) and np.array_equal(H1, H)


def test_full_graph_solve_reduces_error(oracle, small_scene):
    sc = small_scene
    g = oracle.Graph()
    layers = [_layer_of(oracle, s) for s in sc.submaps]
    for i, s in enumerate(sc.submaps):
        g.add_node(i, sc.poses_init[i], constant=(i == 0))
    L = oracle.sqrt_information(sc.odom_information)
    for (i, j, t, y) in sc.odometry:
        g.add_relative(i, j, t, y, L)
    for (i, j) in sc.pairs:
        a, b = sc.submaps[i], sc.submaps[j]
        g.add_registration(i, j, layers[j], a.points_xyz, a.points_distance, a.points_weight)
        g.add_registration(j, i, layers[i], b.points_xyz, b.points_distance, b.points_weight)
    rc, s = g.solve(oracle.solver_options(num_threads=4))
    assert rc == 0 and s.final_cost < s.initial_cost
    e0 = np.abs(sc.poses_init[:, :2] - sc.poses_gt[:, :2]).mean()
    e1 = np.abs(g.poses()[:, :2] - sc.poses_gt[:, :2]).mean()
    assert e1 < e0


def test_lm_minimum_matches_scipy(oracle, pair_scene):
    """Independent optimiser on the same residuals: scipy's trust-region least squares, driven by
    the oracle's residual/Jacobian evaluations, reaches the minimum the oracle's Ceres-style LM
    finds (config-1 problem: odometry edge + mirrored registration constraint)."""
    from scipy.optimize import least_squares
    sc = pair_scene
    s0, s1 = sc.submaps
    layers = [_layer_of(oracle, s) for s in sc.submaps]
    L = oracle.sqrt_information(sc.odom_information)
    (i, j, t_obs, yaw_obs) = sc.odometry[0]
    x0 = sc.poses_init[1].copy(); ref = sc.poses_init[0].copy()

    def fun(x):
        r_rel, _, _ = oracle.relpose_evaluate(ref, x, t_obs, yaw_obs, L)
        _, r01, _, _ = oracle.reg_evaluate(layers[1], s0.points_xyz, s0.points_distance, s0.points_weight,
                                           ref, x, jacobians=False)
        _, r10, _, _ = oracle.reg_evaluate(layers[0], s1.points_xyz, s1.points_distance, s1.points_weight,
                                           x, ref, jacobians=False)
        return np.concatenate([r_rel, r01, r10])

    def jac(x):
        _, _, jb = oracle.relpose_evaluate(ref, x, t_obs, yaw_obs, L)
        _, _, _, je01 = oracle.reg_evaluate(layers[1], s0.points_xyz, s0.points_distance, s0.points_weight, ref, x)
        _, _, jr10, _ = oracle.reg_evaluate(layers[0], s1.points_xyz, s1.points_distance, s1.points_weight, x, ref)
        return np.vstack([jb, je01, jr10])

    res = least_squares(fun, x0, jac=jac, method="trf", xtol=1e-12, ftol=1e-14, gtol=1e-12)
    g = oracle.Graph()
    g.add_node(0, ref, constant=True); g.add_node(1, x0, constant=False)
    g.add_relative(0, 1, t_obs, yaw_obs, L)
    g.add_registration(0, 1, layers[1], s0.points_xyz, s0.points_distance, s0.points_weight)
    g.add_registration(1, 0, layers[0], s1.points_xyz, s1.points_distance, s1.points_weight)
    rc, summ = g.solve(oracle.solver_options(parameter_tolerance=1e-12, function_tolerance=1e-14,
                                             max_num_iterations=200))
    assert rc == 0
    x_lm = g.poses()[1]
    # the cost is piecewise trilinear and evaluated in float32 (kinks + ~1e-7 noise): the two
    # optimisers must agree on the minimum up to that noise floor
    assert abs(summ.final_cost - res.cost) <= 1e-4 * max(res.cost, 1e-9)
    assert np.abs(x_lm - res.x).max() < 1e-3
