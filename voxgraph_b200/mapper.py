"""Host-side mirror of the reference's front end for the streaming path (HP1 -> HP2 online):

  VoxgraphMapper::pointcloudCallback        voxgraph/src/frontend/voxgraph_mapper.cpp:202-265
  VoxgraphMapper::switchToNewSubmap         voxgraph_mapper.cpp:457-510
  VoxgraphMapper::optimizePoseGraph         voxgraph_mapper.cpp:512-524
  PoseGraphInterface::updateRegistrationConstraints / addSubmap / addOdometryMeasurement
                                            src/frontend/pose_graph_interface/pose_graph_interface.cpp:24-175
  VoxgraphSubmapCollection::shouldCreateNewSubmap (submap_creation_interval)

Everything heavy happens behind the C-ABI on the device: the scan is ray-cast into the active
submap (vgx_tsdf_integrate), a finished submap gets its registration view, registration points,
surface OBB and isosurface blocks without leaving the GPU (vgx_submap_finish_ex), overlapping pairs
come from vgx_find_overlapping_pairs and the pose graph is solved by vgx_graph_solve.  ROS, TF,
the odometry map tracker and the visualisation/publishing code of the reference are out of scope:
poses arrive as arguments.
"""
import time
from dataclasses import dataclass, field

import numpy as np

from . import api


def _pose4_to_T(p):
    """[x, y, z, yaw] -> [qw qx qy qz tx ty tz] (minkindr exp of [t, 0, 0, yaw])."""
    p = np.asarray(p, np.float64)
    return np.array([np.cos(0.5 * p[3]), 0.0, 0.0, np.sin(0.5 * p[3]), p[0], p[1], p[2]], np.float32)


def _compose4(a, b):
    """4-DoF pose a composed with b (b expressed in a's frame)."""
    c, s = np.cos(a[3]), np.sin(a[3])
    out = np.array([a[0] + c * b[0] - s * b[1], a[1] + s * b[0] + c * b[1], a[2] + b[2], a[3] + b[3]])
    out[3] = out[3] - 2 * np.pi * np.floor((out[3] + np.pi) / (2 * np.pi))
    return out


def _inverse4(a):
    c, s = np.cos(a[3]), np.sin(a[3])
    return np.array([-(c * a[0] + s * a[1]), -(-s * a[0] + c * a[1]), -a[2], -a[3]])


@dataclass
class MapperConfig:
    voxel_size: float = 0.2                      # voxgraph_mapper.yaml:21
    voxels_per_side: int = 16
    submap_creation_interval: float = 10.0       # seconds (voxgraph_mapper.yaml submap_creation_interval)
    capacity_blocks: int = 16384
    tsdf: dict = field(default_factory=dict)     # overrides of vgx_tsdf_config (mode 1 = Fast by default)
    registration_filter: dict = field(default_factory=dict)
    registration: dict = field(default_factory=dict)   # RegistrationConstraintConfig overrides
    odometry_information: np.ndarray = field(default_factory=lambda: np.diag([1.0, 1.0, 2500.0, 2500.0]))
    registration_constraints_enabled: bool = True
    odometry_constraints_enabled: bool = True


class VoxgraphMapper:
    """Streaming driver: integrate scans, switch submaps on a timer, register online."""

    def __init__(self, ctx, config=None, first_submap_id=0):
        self.ctx = ctx
        self.cfg = config or MapperConfig()
        self.pose_graph = api.PoseGraph(ctx)
        self.submap_ids = []            # creation order
        self.submap_pose = {}           # id -> T_mission_submap [x y z yaw] (optimised)
        self.submap_start = {}
        self._next_id = first_submap_id
        self._last_robot_pose_in_submap = None
        self.timings = []               # one dict per submap switch
        self.scan_stats = []
        kw = dict(mode=1)
        kw.update(self.cfg.tsdf)
        self.tsdf_cfg = ctx.tsdf_config(**kw)
        self.filter = ctx.registration_filter(**self.cfg.registration_filter)

    # ---- VoxgraphSubmapCollection
    def empty(self):
        return not self.submap_ids

    @property
    def active_id(self):
        return self.submap_ids[-1]

    def shouldCreateNewSubmap(self, t):
        return self.empty() or t >= self.submap_start[self.active_id] + self.cfg.submap_creation_interval

    # ---- VoxgraphMapper::pointcloudCallback
    def pointcloudCallback(self, timestamp, T_mission_sensor, points_C):
        """T_mission_sensor: 4-DoF odometry pose of the sensor [x, y, z, yaw]; points_C: (n, 3)."""
        T_mission_sensor = np.asarray(T_mission_sensor, np.float64)
        if self.shouldCreateNewSubmap(timestamp):
            self.switchToNewSubmap(timestamp, T_mission_sensor)
            self.optimizePoseGraph()
        sid = self.active_id
        # map_tracker_.get_T_S_C(): the sensor in the active submap's frame (odometry, not optimised)
        T_S_C = _compose4(_inverse4(self._odom_submap_pose), T_mission_sensor)
        st = self.ctx.tsdf_integrate(sid, _pose4_to_T(T_S_C), points_C, self.tsdf_cfg)
        self._last_T_S_B = T_S_C
        self.scan_stats.append(st)
        return st

    # ---- VoxgraphMapper::switchToNewSubmap
    def switchToNewSubmap(self, timestamp, T_mission_sensor):
        rec = {"finish_ms": 0.0, "overlap_ms": 0.0, "pairs": 0}
        ctx = self.ctx
        if not self.empty():
            t0 = time.time()
            ctx.submap_finish(self.active_id)                          # finishSubmap(): registration view + grid
            ctx.synchronize()
            t1 = time.time()
            ctx.submap_extract_points(self.active_id, self.filter)    # ... registration points, OBB, isosurface blocks
            ctx.synchronize()
            rec["finish_ms"] = (time.time() - t0) * 1e3
            rec["view_ms"] = (t1 - t0) * 1e3
            rec["finished_blocks"] = ctx.submap_block_count(self.active_id)
            rec["isosurface_points"] = ctx.submap_num_points(self.active_id, api.K_ISOSURFACE_POINTS)
        if self.cfg.registration_constraints_enabled and len(self.submap_ids) >= 2:
            t0 = time.time()
            rec["pairs"] = self.updateRegistrationConstraints()
            rec["overlap_ms"] = (time.time() - t0) * 1e3
        # createNewSubmap(T_M_B): the new submap's frame is the robot's current (odometry) pose
        sid = self._next_id
        self._next_id += 1
        ctx.submap_create(sid, self.cfg.voxel_size, self.cfg.voxels_per_side, self.cfg.capacity_blocks)
        prev = self.submap_ids[-1] if self.submap_ids else None
        if prev is None:
            pose = T_mission_sensor.copy()
        else:
            # the optimised pose of the previous submap composed with the odometry since its creation
            T_S1_S2 = _compose4(_inverse4(self._odom_submap_pose), T_mission_sensor)
            pose = _compose4(self.submap_pose[prev], T_S1_S2)
        self.submap_ids.append(sid)
        self.submap_pose[sid] = pose
        self.submap_start[sid] = timestamp
        self.pose_graph.addSubmapNode(api.SubmapNodeConfig(sid, pose, set_constant=(prev is None)))   # addSubmap
        if self.cfg.odometry_constraints_enabled and prev is not None:
            self.pose_graph.addRelativePoseConstraint(api.RelativePoseConstraintConfig(
                prev, sid, T_S1_S2, self.cfg.odometry_information))                                   # addOdometryMeasurement
        self._odom_submap_pose = T_mission_sensor.copy()
        self.timings.append(rec)

    # ---- PoseGraphInterface::updateRegistrationConstraints
    def updateRegistrationConstraints(self):
        pg = self.pose_graph
        pg.resetRegistrationConstraints()
        finished = self.submap_ids        # all of them are finished at this point (the new one does not exist yet)
        T = np.array([_pose4_to_T(self.submap_pose[i]) for i in finished], np.float32)
        pairs = self.ctx.find_overlapping_pairs(finished, T)             # updateOverlappingSubmapList
        for (a, b) in pairs:
            pg.addRegistrationConstraint(api.RegistrationConstraintConfig(a, b, **self.cfg.registration))
        self.overlapping_submap_list = pairs
        return len(pairs)

    # ---- VoxgraphMapper::optimizePoseGraph
    def optimizePoseGraph(self):
        if len(self.submap_ids) < 2:
            return None
        t0 = time.time()
        summ = self.pose_graph.optimize()
        ms = (time.time() - t0) * 1e3
        for i, p in self.pose_graph.getSubmapPoses().items():            # updateSubmapCollectionPoses
            self.submap_pose[i] = p
        self.timings[-1]["optimize_ms"] = ms
        self.timings[-1]["lm_iterations"] = summ.iterations
        self.timings[-1]["registration_blocks"] = len(self.pose_graph.registration_blocks)
        return summ
