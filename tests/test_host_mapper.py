"""Host logic of the streaming front end (voxgraph_b200/mapper.py) without a GPU: a recording stand-in
for the Context checks the order VoxgraphMapper::pointcloudCallback / switchToNewSubmap /
optimizePoseGraph (voxgraph_mapper.cpp:202-265, 457-524) drive the C-ABI in, the frames the scans are
integrated in and the 4-DoF pose algebra."""
import numpy as np

from voxgraph_b200 import mapper as vm


class _Stats:
    voxel_updates = 7


class _Summary:
    iterations = 2


class FakeCtx:
    """Records the calls; graph_solve returns the poses it was given, shifted by +0.5 m in x."""

    def __init__(self):
        self.calls = []
        self.nodes = None
        self.finished = set()
        self.created = []

    # configs
    def tsdf_config(self, **kw):
        return dict(kw)

    def registration_filter(self, **kw):
        return dict(kw)

    def reg_config(self, **kw):
        return dict(kw)

    def solver_options(self, **kw):
        class O:  # noqa
            exclude_registration = 0
        return O()

    # submaps
    def submap_create(self, sid, vs, vps, cap):
        self.created.append(sid)
        self.calls.append(("create", sid, vs, vps, cap))

    def tsdf_integrate(self, sid, T, pts, cfg):
        self.calls.append(("integrate", sid, np.array(T), len(pts)))
        return _Stats()

    def submap_finish(self, sid):
        self.finished.add(sid)
        self.calls.append(("finish", sid))

    def submap_extract_points(self, sid, filt):
        assert sid in self.finished
        self.calls.append(("extract", sid))

    def submap_finish_ex(self, sid, filt):
        self.submap_finish(sid)
        self.submap_extract_points(sid, filt)

    def synchronize(self):
        pass

    def submap_block_count(self, sid):
        return 5

    def submap_num_points(self, sid, kind):
        return 100

    def find_overlapping_pairs(self, ids, T):
        # every submap must have been finished (surface OBB + isosurface blocks exist)
        assert all(i in self.finished for i in ids)
        self.calls.append(("overlap", list(ids), np.array(T)))
        return [(ids[k], ids[k + 1]) for k in range(len(ids) - 1)]

    # pose graph
    def graph_set_nodes(self, ids, x, cst):
        self.nodes = (list(ids), np.array(x, float), list(cst))
        self.calls.append(("nodes", list(ids), list(cst)))

    def graph_set_relative_edges(self, a, b, t, L):
        self.calls.append(("rel", list(a), list(b), np.array(t)))

    def graph_set_registration_constraints_v(self, a, b, cfgs):
        self.calls.append(("reg", list(a), list(b)))

    def graph_solve(self, n, o):
        self.calls.append(("solve", n))
        x = self.nodes[1].copy()
        x[1:, 0] += 0.5      # the "optimiser" moves every free node
        return x, _Summary()


def test_pose_algebra_round_trips():
    rng = np.random.default_rng(3)
    for _ in range(20):
        a = np.array([*rng.uniform(-5, 5, 3), rng.uniform(-3, 3)])
        b = np.array([*rng.uniform(-5, 5, 3), rng.uniform(-3, 3)])
        ident = vm._compose4(a, vm._inverse4(a))
        assert np.allclose(ident, 0, atol=1e-12)
        # (a * b) * b^-1 == a (yaw compared on the circle)
        c = vm._compose4(vm._compose4(a, b), vm._inverse4(b))
        assert np.allclose(c[:3], a[:3], atol=1e-12)
        assert abs(np.angle(np.exp(1j * (c[3] - a[3])))) < 1e-12
    # yaw stays in [-pi, pi)
    w = vm._compose4(np.array([0, 0, 0, 3.0]), np.array([0, 0, 0, 3.0]))
    assert -np.pi <= w[3] < np.pi
    T = vm._pose4_to_T([1.0, 2.0, 3.0, 0.5])
    assert np.allclose(T, [np.cos(0.25), 0, 0, np.sin(0.25), 1, 2, 3])


def test_switch_order_and_frames():
    ctx = FakeCtx()
    cfg = vm.MapperConfig(voxel_size=0.15, submap_creation_interval=1.0, capacity_blocks=64)
    m = vm.VoxgraphMapper(ctx, cfg, first_submap_id=10)
    assert m.empty() and m.shouldCreateNewSubmap(0.0)
    pts = np.zeros((4, 3), np.float32)
    poses = [np.array([0.1 * k, 0.0, 0.0, 0.01 * k]) for k in range(25)]
    for k in range(25):
        m.pointcloudCallback(0.1 * k, poses[k], pts)
    # a new submap every 10 scans: ids 10, 11, 12 created at scans 0, 10, 20
    assert ctx.created == [10, 11, 12] and m.submap_ids == [10, 11, 12]
    assert m.active_id == 12
    # each scan goes into the active submap, in that submap's (odometry) frame: the first scan of a
    # submap is integrated at the identity
    integ = [c for c in ctx.calls if c[0] == "integrate"]
    assert [c[1] for c in integ] == [10] * 10 + [11] * 10 + [12] * 5
    for first in (0, 10, 20):
        assert np.allclose(integ[first][2], [1, 0, 0, 0, 0, 0, 0], atol=1e-7)
    T_5 = vm._pose4_to_T(vm._compose4(vm._inverse4(poses[0]), poses[5]))
    assert np.allclose(integ[5][2], T_5, atol=1e-6)
    # order at a switch: finish(+extract) the old submap -> overlap list over the finished ones ->
    # create the new one -> (nodes, odometry edge, registration blocks) -> solve
    names = [c[0] for c in ctx.calls]
    i_fin = names.index("finish")
    assert ctx.calls[i_fin][1] == 10 and names[i_fin + 1] == "extract"
    i_create2 = [k for k, c in enumerate(ctx.calls) if c[0] == "create"][1]
    assert i_fin < i_create2
    # first switch: one finished submap only -> no overlap query; second switch: both finished submaps
    overlaps = [c for c in ctx.calls if c[0] == "overlap"]
    assert len(overlaps) == 1 and overlaps[0][1] == [10, 11]
    regs = [c for c in ctx.calls if c[0] == "reg"]
    assert regs and regs[-1][1:] == ([10, 11], [11, 10])          # mirrored residual blocks
    rel = [c for c in ctx.calls if c[0] == "rel"][-1]
    assert rel[1] == [10, 11] and rel[2] == [11, 12]              # odometry chain
    # first node constant, the others free
    nodes = [c for c in ctx.calls if c[0] == "nodes"][-1]
    assert nodes[1] == [10, 11, 12] and nodes[2] == [1, 0, 0]
    assert names.count("solve") == 2                               # after the 2nd and 3rd submap exist


def test_optimised_pose_feeds_the_next_submap_origin():
    ctx = FakeCtx()
    m = vm.VoxgraphMapper(ctx, vm.MapperConfig(submap_creation_interval=1.0, capacity_blocks=8))
    pts = np.zeros((1, 3), np.float32)
    for k in range(21):
        m.pointcloudCallback(0.1 * k, np.array([0.1 * k, 0.0, 0.0, 0.0]), pts)
    # the fake optimiser moved submap 1 by +0.5 m at the first solve (then again at the second):
    # submap 2's initial pose = optimised pose of submap 1 composed with the odometry since its creation
    assert np.isclose(m.submap_pose[1][0], 1.0 + 0.5 + 0.5)
    assert m.submap_pose[2][0] >= 2.0 + 0.5          # created from the corrected pose, then moved itself
    assert m.timings[-1]["lm_iterations"] == 2 and m.timings[-1]["registration_blocks"] == 2
