"""Host-side mirror logic (voxgraph_b200/api.py PoseGraph) without a GPU: a recording stand-in
for the Context checks what reaches the C-ABI — mirrored registration blocks, reference-frame
nodes, absolute (height) constraints, sqrt-information factors, CHECK-style errors."""
import numpy as np
import pytest

from voxgraph_b200 import api


class FakeCtx:
    def __init__(self):
        self.calls = []

    def reg_config(self, **kw):
        return dict(kw)

    def solver_options(self, **kw):
        class O:  # noqa
            exclude_registration = 0
        return O()

    def graph_set_nodes(self, ids, x, cst):
        self.calls.append(("nodes", list(ids), np.array(x), list(cst)))

    def graph_set_relative_edges(self, a, b, t, L):
        self.calls.append(("rel", list(a), list(b), np.array(t), np.array(L)))

    def graph_set_registration_constraints(self, a, b, cfg):
        self.calls.append(("reg", list(a), list(b), cfg))

    def graph_set_registration_constraints_v(self, a, b, cfgs):
        # one RegistrationCostFunction::Config per residual block (registration_constraint.h:15-21)
        assert len(cfgs) == len(a) == len(b)
        self.calls.append(("reg", list(a), list(b), cfgs[0] if cfgs else None))

    def graph_solve(self, n, o):
        self.calls.append(("solve", n, o.exclude_registration))
        class S:  # noqa
            iterations = 1
        return np.arange(4 * n, dtype=float).reshape(n, 4), S()


def test_mirroring_and_sync_order():
    ctx = FakeCtx()
    pg = api.PoseGraph(ctx)
    for i in range(3):
        pg.addSubmapNode(api.SubmapNodeConfig(i, np.array([i, 0, 0, 0.1 * i]), set_constant=(i == 0)))
    pg.addRelativePoseConstraint(api.RelativePoseConstraintConfig(0, 1, np.array([1.0, 0, 0, 0.1]),
                                                                  np.diag([1.0, 1.0, 2500.0, 2500.0])))
    pg.addRegistrationConstraint(api.RegistrationConstraintConfig(0, 1))                      # isosurface: mirrored
    pg.addRegistrationConstraint(api.RegistrationConstraintConfig(1, 2, registration_point_type=api.K_VOXELS))
    assert pg.registration_blocks == [(0, 1), (1, 0), (1, 2)]    # pose_graph.cpp:63-71
    s = pg.optimize(exclude_registration_constraints=True)
    kinds = [c[0] for c in ctx.calls]
    assert kinds == ["nodes", "rel", "reg", "solve"]
    assert ctx.calls[0][1] == [0, 1, 2] and ctx.calls[0][3] == [1, 0, 0]
    np.testing.assert_allclose(ctx.calls[1][4][0], np.diag([1.0, 1.0, 50.0, 50.0]))
    assert ctx.calls[2][1] == [0, 1, 1] and ctx.calls[2][2] == [1, 0, 2]
    assert ctx.calls[3][2] == 1
    # poses written back in id order
    np.testing.assert_allclose(pg.getSubmapPoses()[2], [8, 9, 10, 11])
    pg.resetRegistrationConstraints()
    assert pg.registration_blocks == []
    pg.optimize()
    assert [c[0] for c in ctx.calls[4:]] == ["nodes", "rel", "solve"]   # no registration list re-sent


def test_reference_frame_and_height_constraint():
    ctx = FakeCtx()
    pg = api.PoseGraph(ctx)
    pg.addSubmapNode(api.SubmapNodeConfig(0, np.zeros(4), set_constant=True))
    pg.addSubmapNode(api.SubmapNodeConfig(5, np.array([1.0, 2.0, 0.4, 0.0])))
    with pytest.raises(ValueError):
        pg.addAbsolutePoseConstraint(api.AbsolutePoseConstraintConfig(0, 5, np.zeros(4)))   # no frame yet
    pg.addReferenceFrameNode(api.ReferenceFrameNodeConfig(0))
    assert pg.hasReferenceFrameNode(0) and not pg.hasReferenceFrameNode(1)
    info = np.zeros((4, 4)); info[2, 2] = 2500.0            # measurement_templates.cpp:71-80 (height)
    pg.addAbsolutePoseConstraint(api.AbsolutePoseConstraintConfig(
        0, 5, np.array([0, 0, 1.5, 0]), info, allow_semi_definite_information_matrix=True))
    with pytest.raises(ValueError):
        pg.addAbsolutePoseConstraint(api.AbsolutePoseConstraintConfig(0, 5, np.zeros(4), info))  # LLT fails
    pg.optimize()
    nodes = ctx.calls[0]
    assert nodes[1] == [0, 5, api.FRAME_NODE_ID_BASE] and nodes[3] == [1, 0, 1]
    rel = ctx.calls[1]
    assert rel[1] == [api.FRAME_NODE_ID_BASE] and rel[2] == [5]
    np.testing.assert_allclose(rel[4][0], np.diag([0, 0, 50.0, 0]))
    assert set(pg.getSubmapPoses()) == {0, 5}               # frame nodes are not submap poses


def test_sqrt_information_matches_oracle(oracle):
    rs = np.random.RandomState(3)
    for _ in range(100):
        B = rs.normal(size=(4, 4)); info = B @ B.T + 0.1 * np.eye(4)
        np.testing.assert_allclose(api.sqrt_information_matrix(info), oracle.sqrt_information(info), atol=1e-12)
    for _ in range(100):
        B = rs.normal(size=(4, rs.randint(1, 4))); info = B @ B.T
        S = api.sqrt_information_matrix(info, allow_semi_definite=True)
        np.testing.assert_allclose(S @ S.T, info, atol=1e-9)
        np.testing.assert_allclose(S, oracle.sqrt_information_ldlt(info), atol=1e-7)  # rank-deficient: sqrt of 1e-17 round-off
    with pytest.raises(ValueError):
        api.sqrt_information_matrix(-np.eye(4), allow_semi_definite=True)


def test_reference_checks_as_errors():
    pg = api.PoseGraph(FakeCtx())
    pg.addSubmapNode(api.SubmapNodeConfig(0, np.zeros(4), True))
    pg.addSubmapNode(api.SubmapNodeConfig(1, np.zeros(4), False))
    with pytest.raises(ValueError, match="itself"):
        pg.addRegistrationConstraint(api.RegistrationConstraintConfig(0, 0))
    with pytest.raises(ValueError, match="no node"):
        pg.addRegistrationConstraint(api.RegistrationConstraintConfig(0, 7))
    with pytest.raises(ValueError, match="identity"):
        pg.addRegistrationConstraint(api.RegistrationConstraintConfig(0, 1, information_matrix=2 * np.eye(4)))
    with pytest.raises(ValueError, match="positive definite"):
        pg.addRelativePoseConstraint(api.RelativePoseConstraintConfig(0, 1, np.zeros(4), np.zeros((4, 4))))
