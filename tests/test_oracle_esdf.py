"""Known-answer tests of the oracle's ESDF restatement (fixed point of voxblox EsdfIntegrator's batch
update; parity unpinned against the binaries)."""
import numpy as np

from voxgraph_b200 import synth

VS, VPS = 0.2, 16


def test_axis_aligned_plane_esdf_equals_true_distance(oracle):
    """KA: for d = z - c the quasi-Euclidean propagation along the axis is exact: every observed voxel
    carries its true signed distance clipped at +-max_distance; the TSDF band (|d| < 0.2) is copied."""
    idx, d, w = synth.plane_layer_blocks([0, 0, 1.0], 0.37, VS, VPS, ((0, 1), (0, 1), (-1, 1)), trunc=0.6)
    L = oracle.Layer.from_blocks(VS, VPS, idx, d, w)
    e, ob, sweeps = oracle.generate_esdf(L)
    assert ob.min() == 1.0 and 5 <= sweeps <= 40
    lin = np.arange(VPS ** 3)
    for bi, b in enumerate(idx):
        z = (b[2] * VPS + lin // (VPS * VPS) + 0.5) * VS - 0.37
        assert np.abs(e[bi] - np.clip(z, -2.0, 2.0)).max() < 1e-5
    band = np.abs(d) < 0.2
    assert np.array_equal(e[band], d[band])


def test_unobserved_voxels_block_propagation_and_stay_unobserved(oracle):
    idx, d, w = synth.plane_layer_blocks([0, 0, 1.0], 0.37, VS, VPS, ((0, 0), (0, 0), (0, 0)), trunc=0.6)
    w = w.copy().reshape(1, VPS, VPS, VPS)      # [z, y, x]
    w[0, 6, :, :] = 0.0                          # an unobserved slab above the surface
    w = w.reshape(1, -1)
    L = oracle.Layer.from_blocks(VS, VPS, idx, d, w)
    e, ob, _ = oracle.generate_esdf(L)
    e3 = e.reshape(VPS, VPS, VPS); ob3 = ob.reshape(VPS, VPS, VPS)
    assert ob3[6].max() == 0.0 and e3[6].max() == 0.0
    assert np.all(e3[8:] == 2.0)                 # nothing reaches beyond the slab: default distance
    assert np.all(e3[3:6] < 2.0)


def test_sphere_esdf_quasi_euclidean_bounds(oracle):
    """Quasi-Euclidean paths overestimate the Euclidean distance by at most ~8.5 % (26-neighbourhood)."""
    c = np.array([1.7, 1.5, 1.6]); R = 0.9
    idx, d, w = synth.field_layer_blocks(lambda p: np.linalg.norm(p - c, axis=-1) - R, VS, VPS,
                                         ((0, 1), (0, 1), (0, 1)), weight=1.0)
    tsdf = np.clip(d, -0.6, 0.6).astype(np.float32)
    L = oracle.Layer.from_blocks(VS, VPS, idx, tsdf, w)
    e, ob, _ = oracle.generate_esdf(L)
    true = np.clip(d, -2.0, 2.0)
    m = (np.abs(true) < 1.9) & (np.abs(true) > 0.2)
    err = e[m] - true[m]
    assert np.all(np.sign(e[m]) == np.sign(true[m]))
    assert np.abs(err).max() < 0.1 * np.abs(true[m]).max() + VS
