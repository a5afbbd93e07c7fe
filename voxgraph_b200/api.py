"""Host-side mirror of the reference's interface for the two hot paths, on top of the C-ABI.

Names, argument meaning and error behaviour follow the reference:
  PoseGraph                 voxgraph/include/voxgraph/backend/pose_graph.h:15-65
  RegistrationCostFunction  .../cost_functions/registration_cost_function.h:11-82
  PointcloudIntegrator      voxgraph/include/voxgraph/frontend/measurement_processors/pointcloud_integrator.h
(the C++ mirror of the same surface lives in voxgraph_b200/host/; this module is what the
pytest parity tests and bench.py drive).
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import (EsdfConfig, RegConfig, RegistrationFilter, SolverOptions, SolverSummary, TsdfConfig, TsdfStats)

K_VOXELS = 0            # VoxgraphSubmap::RegistrationPointType::kVoxels
K_ISOSURFACE_POINTS = 1  # ...::kIsosurfacePoints


class VgxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vgx error %d: %s" % (code, msg))
        self.code = code


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One GPU's brick store + pose graph state (vgx_ctx)."""

    def __init__(self, device=0):
        self._L = _lib.load()
        h = C.c_void_p()
        rc = self._L.vgx_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise VgxError(rc, "vgx_ctx_create failed (a CUDA device is required; no CPU fallback)")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.vgx_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow=(0,)):
        if rc not in allow:
            raise VgxError(rc, self._L.vgx_last_error(self._h).decode())
        return rc

    # ---- stream / accounting
    @property
    def stream_ptr(self):
        return self._L.vgx_ctx_stream(self._h)

    def synchronize(self):
        self._check(self._L.vgx_ctx_synchronize(self._h))

    def profile_enable(self, on=True):
        self._check(self._L.vgx_profile_enable(self._h, int(on)))

    def profile_reset(self):
        self._check(self._L.vgx_profile_reset(self._h))

    def profile_get(self, which):
        ms = C.c_double(0); n = C.c_int64(0)
        self._check(self._L.vgx_profile_get(self._h, which, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    @property
    def launch_count(self):
        return self._L.vgx_launch_count(self._h)

    # ---- submaps
    def submap_upload(self, submap_id, voxel_size, vps, block_idx, distance, weight):
        bi = np.ascontiguousarray(block_idx, dtype=np.int32).reshape(-1, 3)
        n = bi.shape[0]
        d = _f32(distance).reshape(n, -1); w = _f32(weight).reshape(n, -1)
        assert n == 0 or (d.shape[1] == vps ** 3 and w.shape == d.shape)
        self._check(self._L.vgx_submap_upload(self._h, int(submap_id), float(voxel_size), int(vps), n,
                                              _p(bi, C.c_int32), _p(d, C.c_float), _p(w, C.c_float)))

    def submap_create(self, submap_id, voxel_size, vps, capacity_blocks):
        self._check(self._L.vgx_submap_create(self._h, int(submap_id), float(voxel_size), int(vps),
                                              int(capacity_blocks)))

    def submap_finish(self, submap_id):
        self._check(self._L.vgx_submap_finish(self._h, int(submap_id)))

    def submap_peek_device(self, submap_id):
        """(device pointer of block indices, device pointer of (distance, weight) voxels, n_blocks)."""
        pi = C.c_void_p(); pd = C.c_void_p(); n = C.c_int(0)
        self._check(self._L.vgx_submap_peek_device(self._h, int(submap_id), C.byref(pi), C.byref(pd), C.byref(n)))
        return pi.value, pd.value, n.value

    def submap_upload_device(self, submap_id, voxel_size, vps, n_blocks, d_block_idx_ptr, d_dw_ptr):
        self._check(self._L.vgx_submap_upload_device(self._h, int(submap_id), float(voxel_size), int(vps),
                                                     int(n_blocks), C.c_void_p(d_block_idx_ptr), C.c_void_p(d_dw_ptr)))

    def registration_filter(self, **kw):
        f = RegistrationFilter()
        self._L.vgx_registration_filter_default(C.byref(f))
        for k, v in kw.items():
            setattr(f, k, v)
        return f

    def submap_extract_points(self, submap_id, filt=None):
        """findRelevantVoxelIndices + findIsosurfaceVertices + surface OBB on the device."""
        self._check(self._L.vgx_submap_extract_points(self._h, int(submap_id),
                                                      C.byref(filt) if filt is not None else None))

    def submap_finish_ex(self, submap_id, filt=None):
        self._check(self._L.vgx_submap_finish_ex(self._h, int(submap_id),
                                                 C.byref(filt) if filt is not None else None))

    def submap_num_points(self, submap_id, point_type):
        n = C.c_int(0)
        self._check(self._L.vgx_submap_num_points(self._h, int(submap_id), int(point_type), C.byref(n)))
        return n.value

    def submap_download_points(self, submap_id, point_type):
        n = self.submap_num_points(submap_id, point_type)
        xyz = np.zeros((n, 3), np.float32); d = np.zeros(n, np.float32); w = np.zeros(n, np.float32)
        got = C.c_int(0)
        self._check(self._L.vgx_submap_download_points(self._h, int(submap_id), int(point_type), n,
                                                       _p(xyz, C.c_float), _p(d, C.c_float),
                                                       _p(w, C.c_float), C.byref(got)))
        return xyz, d, w

    def submap_surface_obb(self, submap_id):
        mn = np.zeros(3, np.float32); mx = np.zeros(3, np.float32)
        rc = self._check(self._L.vgx_submap_surface_obb(self._h, int(submap_id), _p(mn, C.c_float),
                                                        _p(mx, C.c_float)), allow=(0, 1))
        return rc == 0, mn, mx

    def esdf_config(self, **kw):
        c = EsdfConfig()
        self._L.vgx_esdf_config_default(C.byref(c))
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    def submap_generate_esdf(self, submap_id, cfg=None):
        """TsdfEsdfSubmap::generateEsdf on the device -> number of relaxation sweeps."""
        n = C.c_int(0)
        self._check(self._L.vgx_submap_generate_esdf(self._h, int(submap_id),
                                                     C.byref(cfg) if cfg is not None else None, C.byref(n)))
        return n.value

    def submap_download_esdf(self, submap_id):
        _, vps, n, _ = self.submap_info(submap_id)
        d = np.zeros((n, vps ** 3), np.float32); ob = np.zeros((n, vps ** 3), np.float32)
        got = C.c_int(0)
        self._check(self._L.vgx_submap_download_esdf(self._h, int(submap_id), n, _p(d, C.c_float),
                                                     _p(ob, C.c_float), C.byref(got)))
        return d, ob

    def find_overlapping_pairs(self, submap_ids, poses_T):
        """updateOverlappingSubmapList: poses_T (n,7) = [qw qx qy qz tx ty tz] -> list of (id_i, id_j)."""
        ids = np.ascontiguousarray(submap_ids, np.uint32)
        T = _f32(poses_T).reshape(-1, 7)
        assert T.shape[0] == len(ids)
        cap = max(1, len(ids) * (len(ids) - 1) // 2)
        out = np.zeros((cap, 2), np.uint32); n = C.c_int(0)
        self._check(self._L.vgx_find_overlapping_pairs(self._h, len(ids), _p(ids, C.c_uint32),
                                                       _p(T, C.c_float), cap, _p(out, C.c_uint32),
                                                       C.byref(n)))
        return [tuple(int(v) for v in out[k]) for k in range(n.value)]

    def submap_free(self, submap_id):
        self._check(self._L.vgx_submap_free(self._h, int(submap_id)))

    def submap_block_count(self, submap_id):
        n = C.c_int(0)
        self._check(self._L.vgx_submap_block_count(self._h, int(submap_id), C.byref(n)))
        return n.value

    def submap_info(self, submap_id):
        """(voxel_size, voxels_per_side, n_blocks, finished)"""
        vs = C.c_float(0); vps = C.c_int(0); n = C.c_int(0); fin = C.c_int(0)
        self._check(self._L.vgx_submap_info(self._h, int(submap_id), C.byref(vs), C.byref(vps),
                                            C.byref(n), C.byref(fin)))
        return vs.value, vps.value, n.value, bool(fin.value)

    def submap_draw_samples(self, submap_id, point_type, n):
        """WeightedSampler::getRandomItem n times on the submap's generator -> indices."""
        idx = np.zeros(n, np.int32)
        self._check(self._L.vgx_submap_draw_samples(self._h, int(submap_id), int(point_type), int(n),
                                                    _p(idx, C.c_int32)))
        return idx

    def submap_download(self, submap_id, vps=None):
        _, vps_dev, n, _ = self.submap_info(submap_id)
        vps = vps_dev   # buffers are sized from the layer's own voxels_per_side
        idx = np.zeros((n, 3), np.int32)
        d = np.zeros((n, vps ** 3), np.float32); w = np.zeros((n, vps ** 3), np.float32)
        got = C.c_int(0)
        self._check(self._L.vgx_submap_download(self._h, int(submap_id), n, _p(idx, C.c_int32),
                                                _p(d, C.c_float), _p(w, C.c_float), C.byref(got)))
        return idx, d, w

    def submap_upload_points(self, submap_id, point_type, xyz, distance, weight):
        xyz = _f32(xyz).reshape(-1, 3); distance = _f32(distance); weight = _f32(weight)
        self._check(self._L.vgx_submap_upload_points(self._h, int(submap_id), int(point_type),
                                                     xyz.shape[0], _p(xyz, C.c_float),
                                                     _p(distance, C.c_float), _p(weight, C.c_float)))

    def upload_synth_submap(self, s, point_type=K_ISOSURFACE_POINTS):
        """Upload a voxgraph_b200.synth.Submap (bricks + registration points)."""
        self.submap_upload(s.submap_id, s.voxel_size, s.vps, s.block_idx, s.distance, s.weight)
        if s.points_xyz is not None:
            self.submap_upload_points(s.submap_id, point_type, s.points_xyz, s.points_distance,
                                      s.points_weight)

    # ---- b3
    def tsdf_config(self, **kw):
        c = TsdfConfig()
        self._L.vgx_tsdf_config_default(C.byref(c))
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    def tsdf_integrate(self, submap_id, T_G_C, points_C, cfg=None, rgba=None):
        T = _f32(T_G_C); pts = _f32(points_C).reshape(-1, 3)
        st = TsdfStats()
        col = np.ascontiguousarray(rgba, np.uint8) if rgba is not None else None
        self._check(self._L.vgx_tsdf_integrate(self._h, int(submap_id), _p(T, C.c_float), pts.shape[0],
                                               _p(pts, C.c_float), _p(col, C.c_uint8),
                                               C.byref(cfg) if cfg is not None else None,
                                               C.byref(st)))
        return st

    # ---- b1
    def reg_config(self, **kw):
        c = RegConfig()
        self._L.vgx_reg_config_default(C.byref(c))
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    def reg_num_residuals(self, ref_id, cfg=None):
        cfg = cfg or self.reg_config()
        n = C.c_int(0)
        self._check(self._L.vgx_reg_num_residuals(self._h, int(ref_id), C.byref(cfg), C.byref(n)))
        return n.value

    def reg_eval_emit(self, ref_id, read_id, ref_pose, read_pose, cfg=None, jacobians=True):
        """RegistrationCostFunction::Evaluate -> (ok, residuals, J_ref, J_read)."""
        cfg = cfg or self.reg_config()
        K = self.reg_num_residuals(ref_id, cfg)
        r = np.zeros(K); jr = np.zeros((K, 4)) if jacobians else None
        je = np.zeros((K, 4)) if jacobians else None
        a = _f64(ref_pose); b = _f64(read_pose)
        rc = self._check(self._L.vgx_reg_eval_emit(self._h, int(ref_id), int(read_id), C.byref(cfg),
                                                   _p(a, C.c_double), _p(b, C.c_double),
                                                   _p(r, C.c_double), _p(jr, C.c_double),
                                                   _p(je, C.c_double)), allow=(0, 1))
        return rc == 0, r, jr, je

    # ---- b2 (raw)
    def graph_set_nodes(self, ids, xyzyaw, constant):
        ids = np.ascontiguousarray(ids, np.uint32); x = _f64(xyzyaw).reshape(-1, 4)
        cst = np.ascontiguousarray(constant, np.uint8)
        self._check(self._L.vgx_graph_set_nodes(self._h, len(ids), _p(ids, C.c_uint32),
                                                _p(x, C.c_double), _p(cst, C.c_uint8)))

    def graph_set_poses(self, xyzyaw):
        x = _f64(xyzyaw)
        self._check(self._L.vgx_graph_set_poses(self._h, _p(x, C.c_double)))

    def graph_get_poses(self, n):
        x = np.zeros((n, 4))
        self._check(self._L.vgx_graph_get_poses(self._h, _p(x, C.c_double)))
        return x

    def graph_set_relative_edges(self, ids_a, ids_b, t_obs_xyzyaw, sqrt_info):
        a = np.ascontiguousarray(ids_a, np.uint32); b = np.ascontiguousarray(ids_b, np.uint32)
        t = _f64(t_obs_xyzyaw).reshape(-1, 4); L = _f64(sqrt_info).reshape(-1, 16)
        self._check(self._L.vgx_graph_set_relative_edges(self._h, len(a), _p(a, C.c_uint32),
                                                         _p(b, C.c_uint32), _p(t, C.c_double),
                                                         _p(L, C.c_double)))

    def graph_set_registration_constraints(self, ref_ids, read_ids, cfg=None):
        a = np.ascontiguousarray(ref_ids, np.uint32); b = np.ascontiguousarray(read_ids, np.uint32)
        cfg = cfg or self.reg_config()
        self._check(self._L.vgx_graph_set_registration_constraints(self._h, len(a), _p(a, C.c_uint32),
                                                                   _p(b, C.c_uint32), C.byref(cfg)))

    def graph_set_registration_constraints_v(self, ref_ids, read_ids, cfgs):
        a = np.ascontiguousarray(ref_ids, np.uint32); b = np.ascontiguousarray(read_ids, np.uint32)
        arr = (RegConfig * len(a))(*cfgs)
        self._check(self._L.vgx_graph_set_registration_constraints_v(self._h, len(a), _p(a, C.c_uint32),
                                                                     _p(b, C.c_uint32), arr))

    def graph_get_sample_indices(self, constraint, max_n):
        idx = np.zeros(max(max_n, 1), np.int32); n = C.c_int(0)
        self._check(self._L.vgx_graph_get_sample_indices(self._h, int(constraint), int(max_n),
                                                         _p(idx, C.c_int32), C.byref(n)))
        return idx[:n.value].copy()

    def graph_set_sample_indices(self, constraint, indices):
        idx = np.ascontiguousarray(indices, np.int32)
        self._check(self._L.vgx_graph_set_sample_indices(self._h, int(constraint), len(idx),
                                                         _p(idx, C.c_int32)))

    def graph_num_registration_residuals(self):
        a = C.c_int64(0); b = C.c_int64(0)
        self._check(self._L.vgx_graph_num_registration_residuals(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def graph_eval(self, n_nodes, exclude_registration=False, want_H=True):
        dim = 4 * n_nodes
        cost = C.c_double(0); g = np.zeros(dim); H = np.zeros((dim, dim)) if want_H else None
        rc = self._check(self._L.vgx_graph_eval(self._h, int(exclude_registration), C.byref(cost),
                                                _p(g, C.c_double), _p(H, C.c_double)), allow=(0, 1))
        return rc == 0, cost.value, g, H

    def graph_eval_async(self, exclude_registration=False):
        self._check(self._L.vgx_graph_eval_async(self._h, int(exclude_registration)))

    def graph_registration_costs(self, n_constraints):
        out = np.zeros(n_constraints)
        self._check(self._L.vgx_graph_registration_costs(self._h, _p(out, C.c_double)))
        return out

    def graph_edge_covariances(self, ids_a, ids_b):
        a = np.ascontiguousarray(ids_a, np.uint32); b = np.ascontiguousarray(ids_b, np.uint32)
        cov = np.zeros((len(a), 4, 4))
        self._check(self._L.vgx_graph_edge_covariances(self._h, len(a), _p(a, C.c_uint32), _p(b, C.c_uint32),
                                                       _p(cov, C.c_double)))
        return cov

    def solver_options(self, **kw):
        o = SolverOptions()
        self._L.vgx_solver_options_default(C.byref(o))
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def graph_solve(self, n_nodes, options=None):
        o = options or self.solver_options()
        x = np.zeros((n_nodes, 4)); s = SolverSummary()
        self._check(self._L.vgx_graph_solve(self._h, C.byref(o), _p(x, C.c_double), C.byref(s)))
        return x, s

    # ---- multi-GPU
    def comm_suspend(self, on=True):
        """Behave as a single-rank context (full problem locally, no exchange) while on."""
        self._check(self._L.vgx_comm_suspend(self._h, int(bool(on))))

    def comm_init(self, nranks, rank, unique_id):
        u = np.ascontiguousarray(unique_id, np.uint8)
        assert u.size == 128
        self._check(self._L.vgx_comm_init(self._h, int(nranks), int(rank), _p(u, C.c_uint8)))


def p2p_setup(ctx, nranks, rank, all_gather_bytes, capacity_doubles=1 << 20):
    """Bring up the NVLink peer exchange: export this rank's buffer, all-gather the 64-byte IPC
    handles with the caller's collective (all_gather_bytes(np.uint8[64]) -> np.uint8[nranks, 64]),
    map the peers."""
    h = ctx.comm_p2p_export(capacity_doubles)
    handles = np.ascontiguousarray(all_gather_bytes(h), np.uint8).reshape(nranks, 64)
    ctx.comm_p2p_import(nranks, rank, handles)


def shard_constraints(nranks, num_residuals, locality_keys=None):
    """Constraint -> rank partition used by the library (host only, no GPU needed)."""
    cnt = np.ascontiguousarray(num_residuals, np.int32)
    owner = np.zeros(len(cnt), np.int32)
    keys = np.ascontiguousarray(locality_keys, np.uint32) if locality_keys is not None else None
    rc = _lib.load().vgx_shard_constraints(int(nranks), len(cnt), _p(cnt, C.c_int32),
                                           _p(keys, C.c_uint32), _p(owner, C.c_int32))
    if rc != 0:
        raise VgxError(rc, "vgx_shard_constraints failed")
    return owner


def _ctx_p2p_export(self, capacity_doubles):
    h = np.zeros(64, np.uint8)
    self._check(self._L.vgx_comm_p2p_export(self._h, int(capacity_doubles), _p(h, C.c_uint8)))
    return h


def _ctx_p2p_import(self, nranks, rank, handles):
    hs = np.ascontiguousarray(handles, np.uint8).reshape(-1)
    assert hs.size == 64 * nranks
    self._check(self._L.vgx_comm_p2p_import(self._h, int(nranks), int(rank), _p(hs, C.c_uint8)))


Context.comm_p2p_export = _ctx_p2p_export
Context.comm_p2p_import = _ctx_p2p_import


def comm_unique_id():
    u = np.zeros(128, np.uint8)
    rc = _lib.load().vgx_comm_unique_id(_p(u, C.c_uint8))
    if rc != 0:
        raise VgxError(rc, "vgx_comm_unique_id failed")
    return u


# --------------------------------------------------------------------------- PoseGraph mirror
@dataclass
class SubmapNodeConfig:        # SubmapNode::Config (submap_node.h:16-18, node.h:17-21)
    submap_id: int
    T_mission_node_initial: np.ndarray  # [x, y, z, yaw] (Pose4D, pose_4d.cpp:4-11)
    set_constant: bool = False


@dataclass
class RelativePoseConstraintConfig:   # relative_pose_constraint.h:13-18
    origin_submap_id: int
    destination_submap_id: int
    T_origin_destination: np.ndarray  # [tx, ty, tz, yaw]
    information_matrix: np.ndarray = field(default_factory=lambda: np.eye(4))


@dataclass
class ReferenceFrameNodeConfig:       # ReferenceFrameNode::Config (reference_frame_node.h:15-17)
    reference_frame_id: int
    T_mission_node_initial: np.ndarray = field(default_factory=lambda: np.zeros(4))
    set_constant: bool = True


@dataclass
class AbsolutePoseConstraintConfig:   # absolute_pose_constraint.h:13-18
    reference_frame_id: int
    submap_id: int
    T_ref_submap: np.ndarray          # [tx, ty, tz, yaw]
    information_matrix: np.ndarray = field(default_factory=lambda: np.eye(4))
    allow_semi_definite_information_matrix: bool = False


def sqrt_information_matrix(information_matrix, allow_semi_definite=False):
    """Constraint ctor (constraint.cpp:4-38): the LLT lower factor, or for semi-definite matrices
    Eigen's pivoted LDLT, sqrt = P^T L sqrt(D) P. Raises like the reference CHECKs."""
    info = np.asarray(information_matrix, np.float64).reshape(4, 4)
    if not allow_semi_definite:
        try:
            return np.linalg.cholesky(info)
        except np.linalg.LinAlgError:
            raise ValueError("The square root of the information matrix could not be computed, "
                             "make sure it is symmetric and positive definite")
    A = info.copy()
    perm = np.arange(4)
    for k in range(4):
        piv = k + int(np.argmax(np.abs(np.diag(A)[k:])))
        if piv != k:
            A[[k, piv], :] = A[[piv, k], :]
            A[:, [k, piv]] = A[:, [piv, k]]
            perm[[k, piv]] = perm[[piv, k]]
        d = A[k, k]
        if d < -1e-12 * (1.0 + np.abs(np.diag(info)).max()):
            raise ValueError("The information matrix must be positive semi-definite")
        if abs(d) > 0:
            A[k + 1:, k] /= d
            A[k + 1:, k + 1:] -= np.outer(A[k + 1:, k], A[k + 1:, k]) * d
        else:
            A[k + 1:, k] = 0
        A[k, k + 1:] = 0
    Lm = np.tril(A, -1) + np.eye(4)
    M = Lm * np.sqrt(np.maximum(np.diag(A), 0.0))[None, :]
    S = np.zeros((4, 4))
    S[np.ix_(perm, perm)] = M
    return S


FRAME_NODE_ID_BASE = 0x80000000   # reference-frame nodes share the C-ABI's uint32 node id space


@dataclass
class RegistrationConstraintConfig:   # registration_constraint.h:15-21
    first_submap_id: int
    second_submap_id: int
    registration_point_type: int = K_ISOSURFACE_POINTS
    no_correspondence_cost: float = 0.0
    sampling_ratio: float = -1.0
    information_matrix: np.ndarray = field(default_factory=lambda: np.eye(4))
    use_esdf_distance: bool = False       # registration_cost_function.h:35 (reference default true)


class PoseGraph:
    """voxgraph::PoseGraph over the GPU path (pose_graph.h:15-65). Submaps are referred to by id
    and must have been uploaded to the Context (bricks + registration points)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self._nodes = {}
        self._frames = set()
        self._relative = []
        self._registration = []   # (ref_id, read_id) residual blocks, mirrored ones included
        self._reg_cfgs = []       # one RegistrationCostFunction::Config per residual block
        self._dirty = True
        self.solver_summaries = []
        self.solver_options = ctx.solver_options()   # pose_graph.cpp:91-97 defaults

    def addSubmapNode(self, config):
        self._nodes[int(config.submap_id)] = [np.asarray(config.T_mission_node_initial, np.float64).copy(),
                                              bool(config.set_constant)]
        self._dirty = True

    def hasSubmapNode(self, submap_id):
        return int(submap_id) in self._nodes

    def addReferenceFrameNode(self, config):
        """pose_graph.cpp:21-24; the frame node lives at id FRAME_NODE_ID_BASE + frame id."""
        nid = FRAME_NODE_ID_BASE + int(config.reference_frame_id)
        self._nodes[nid] = [np.asarray(config.T_mission_node_initial, np.float64).copy(),
                            bool(config.set_constant)]
        self._frames.add(int(config.reference_frame_id))
        self._dirty = True

    def hasReferenceFrameNode(self, frame_id):
        return int(frame_id) in self._frames

    def addAbsolutePoseConstraint(self, config):
        """pose_graph.cpp:33-39 + absolute_pose_constraint.cpp:6-35: a RelativePoseCostFunction between
        the reference-frame node and the submap node (e.g. the height constraint, information zz only)."""
        if int(config.reference_frame_id) not in self._frames:
            raise ValueError("Graph contains no reference frame node %d" % int(config.reference_frame_id))
        if int(config.submap_id) not in self._nodes:
            raise ValueError("Graph contains no node for submap %d" % int(config.submap_id))
        L = sqrt_information_matrix(config.information_matrix,
                                    config.allow_semi_definite_information_matrix)
        self._relative.append((FRAME_NODE_ID_BASE + int(config.reference_frame_id), int(config.submap_id),
                               np.asarray(config.T_ref_submap, np.float64).copy(), L))
        self._dirty = True

    def addRelativePoseConstraint(self, config):
        # Constraint ctor (constraint.cpp:8-14): LLT lower factor, CHECK on failure
        L = sqrt_information_matrix(config.information_matrix,
                                    getattr(config, "allow_semi_definite_information_matrix", False))
        self._relative.append((int(config.origin_submap_id), int(config.destination_submap_id),
                               np.asarray(config.T_origin_destination, np.float64).copy(), L))
        self._dirty = True

    def addRegistrationConstraint(self, config):
        a, b = int(config.first_submap_id), int(config.second_submap_id)
        if a == b:   # pose_graph.cpp:50-51
            raise ValueError("Cannot constrain submap %d to itself" % a)
        for s in (a, b):   # pose_graph.cpp:54-57
            if s not in self._nodes:
                raise ValueError("Graph contains no node for submap %d" % s)
        if not np.array_equal(np.asarray(config.information_matrix), np.eye(4)):
            raise ValueError("Registration constraint information matrices that differ from the "
                             "identity matrix are not yet supported.")   # registration_constraint.h:32-34
        cfg = self.ctx.reg_config(registration_point_type=int(config.registration_point_type),
                                  no_correspondence_cost=float(config.no_correspondence_cost),
                                  sampling_ratio=float(config.sampling_ratio),
                                  use_esdf_distance=int(bool(getattr(config, "use_esdf_distance", False))))
        self._registration.append((a, b)); self._reg_cfgs.append(cfg)
        if config.registration_point_type == K_ISOSURFACE_POINTS:   # pose_graph.cpp:63-71
            self._registration.append((b, a)); self._reg_cfgs.append(cfg)
        self._dirty = True

    def resetRegistrationConstraints(self):
        self._registration = []
        self._reg_cfgs = []
        self._dirty = True

    def _sync(self):
        ids = sorted(self._nodes)
        if self._dirty:
            x = np.array([self._nodes[i][0] for i in ids]).reshape(-1, 4)
            cst = np.array([self._nodes[i][1] for i in ids], np.uint8)
            self.ctx.graph_set_nodes(ids, x, cst)
            if self._relative:
                self.ctx.graph_set_relative_edges([r[0] for r in self._relative],
                                                  [r[1] for r in self._relative],
                                                  np.array([r[2] for r in self._relative]),
                                                  np.array([r[3] for r in self._relative]))
            if self._registration:
                self.ctx.graph_set_registration_constraints_v([r[0] for r in self._registration],
                                                              [r[1] for r in self._registration],
                                                              self._reg_cfgs)
            self._dirty = False
        return ids

    def optimize(self, exclude_registration_constraints=False):
        ids = self._sync()
        o = self.solver_options
        o.exclude_registration = int(exclude_registration_constraints)
        x, s = self.ctx.graph_solve(len(ids), o)
        for k, i in enumerate(ids):
            self._nodes[i][0] = x[k].copy()
        self.solver_summaries.append(s)
        return s

    def evaluate(self, exclude_registration_constraints=False, want_H=True):
        ids = self._sync()
        return self.ctx.graph_eval(len(ids), exclude_registration_constraints, want_H)

    def getSubmapPoses(self):
        return {i: v[0].copy() for i, v in self._nodes.items() if i < FRAME_NODE_ID_BASE}

    def getEdgeCovarianceMap(self, submap_id_pairs):
        """pose_graph.cpp:117-163: {(first, second): 4x4 covariance} for the requested pairs."""
        self._sync()
        pairs = [(int(a), int(b)) for (a, b) in submap_id_pairs]
        cov = self.ctx.graph_edge_covariances([p[0] for p in pairs], [p[1] for p in pairs])
        return {p: cov[k] for k, p in enumerate(pairs)}

    def getSolverSummaries(self):
        return self.solver_summaries

    def getVisualizationEdgeResiduals(self):
        """Summed squared residual per registration residual block (pose_graph.cpp:194-207)."""
        self._sync()
        return self.ctx.graph_registration_costs(len(self._registration))

    @property
    def registration_blocks(self):
        return list(self._registration)
