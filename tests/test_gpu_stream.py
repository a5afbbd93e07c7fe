"""Streaming path (BASELINE configs[2] shape, scaled down): scans are integrated into the active
submap, a finished submap gets its registration points on the device, overlapping pairs and the
pose-graph solve follow - VoxgraphMapper::pointcloudCallback's order - with no voxel ever
downloaded between HP1 and HP2."""
import numpy as np
import pytest

from voxgraph_b200 import synth

pytestmark = pytest.mark.gpu


def test_integrate_finish_register_online():
    from voxgraph_b200 import api, mapper as vm
    ctx = api.Context(0)
    world = synth.make_world(7, size_xy=(40.0, 30.0), n_clutter=60, n_walls=5)
    n = 16
    s_arc = np.arange(n) * 0.6
    gt = np.stack([8.0 + s_arc, 12.0 + 0.5 * np.sin(s_arc / 3.0), np.full(n, 1.2), 0.05 * s_arc], -1)
    rng = np.random.default_rng(0)
    odo = gt + np.cumsum(rng.normal(0, 1.0, (n, 4)) * np.array([0.02, 0.02, 0.001, 0.002]), 0)
    scans = [synth.lidar_scan(world, gt[k], n_beams=24, n_azimuth=360, seed=k, max_range=12.0) for k in range(n)]
    cfg = vm.MapperConfig(voxel_size=0.2, submap_creation_interval=4 / 10.0, capacity_blocks=4096,
                          tsdf=dict(max_ray_length_m=12.0))
    m = vm.VoxgraphMapper(ctx, cfg)
    launches0 = ctx.launch_count
    for k in range(n):
        st = m.pointcloudCallback(k / 10.0, odo[k], scans[k])
        assert st.voxel_updates > 0
    assert len(m.submap_ids) == 4
    done = [t for t in m.timings if t.get("finish_ms", 0) > 0]
    assert len(done) == 3
    for t in done:
        assert t["isosurface_points"] > 200 and t["finished_blocks"] > 10
    # the last switch saw three finished submaps that overlap pairwise along the path
    assert done[-1]["pairs"] >= 2 and done[-1]["registration_blocks"] == 2 * done[-1]["pairs"]
    assert "optimize_ms" in done[-1] and done[-1]["lm_iterations"] >= 1
    # finished submaps are registration-ready on the device: evaluate a constraint directly
    a, b = m.overlapping_submap_list[0]
    ok, r, jr, je = ctx.reg_eval_emit(a, b, m.submap_pose[a], m.submap_pose[b])
    assert ok and np.isfinite(r).all() and (np.abs(jr).sum(1) > 0).sum() > 50
    # optimisation keeps the chain close to the ground truth (gauge: first submap)
    starts = [int(round(m.submap_start[i] * 10.0)) for i in m.submap_ids]
    opt = np.array([m.submap_pose[i] for i in m.submap_ids])
    err_opt = np.abs((opt[:, :2] - opt[0, :2]) - (gt[starts][:, :2] - gt[starts][0, :2])).max()
    assert err_opt < 0.5
    assert ctx.launch_count > launches0
    ctx.close()
