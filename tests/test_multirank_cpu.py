"""N > 1 host logic on CPU: the library's constraint partition + 'sum of per-rank packed normal
equations == global normal equations', over a real 2-process gloo group."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["VGX_ROOT"])
from voxgraph_b200 import api, synth
from oracle import oracle as o
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sc = synth.make_scene(seed=2, n_submaps=6, n_points=1500, radius=8.0, size_xy=(48.0, 32.0),
                      n_clutter=80, n_walls=6)
blocks = []
for (i, j) in sc.pairs:
    blocks += [(i, j), (j, i)]
counts = [sc.submaps[a].points_xyz.shape[0] - 37 * (k % 3) for k, (a, b) in enumerate(blocks)]
owner = api.shard_constraints(world, counts, [b for (a, b) in blocks])
layers = [o.Layer.from_blocks(s.voxel_size, s.vps, s.block_idx, s.distance, s.weight) for s in sc.submaps]
L = o.sqrt_information(sc.odom_information)

def graph(keep):
    g = o.Graph()
    for i in range(len(sc.submaps)):
        g.add_node(i, sc.poses_init[i], constant=(i == 0))
    if keep is None or rank == 0:          # relative-pose blocks live on rank 0
        for (i, j, t, y) in sc.odometry:
            g.add_relative(i, j, t, y, L)
    for k, (a, b) in enumerate(blocks):
        if keep is None or owner[k] == keep:
            n = counts[k]
            s = sc.submaps[a]
            g.add_registration(a, b, layers[b], s.points_xyz[:n], s.points_distance[:n], s.points_weight[:n])
    return g

ok, cost, grad, H = graph(rank).eval()
t = torch.from_numpy(np.concatenate([[cost], grad, H.ravel()]))
dist.all_reduce(t)                         # the one exchange of the path
ok, cost_f, grad_f, H_f = graph(None).eval()
full = np.concatenate([[cost_f], grad_f, H_f.ravel()])
err = np.abs(t.numpy() - full).max() / np.abs(full).max()
loads = np.bincount(owner, weights=counts, minlength=world)
assert err < 1e-12, err
assert loads.max() - loads.min() <= max(counts), loads
assert sorted(set(owner.tolist())) == list(range(world))
sys.stdout.write("rank-%d-ok %g %s\n" % (rank, err, loads.tolist())); sys.stdout.flush()
dist.destroy_process_group()
'''


def test_two_rank_gloo_partition_and_sum(tmp_path):
    from voxgraph_b200 import build
    from oracle import oracle as o
    build.build(); o.build()      # build once here, not concurrently in the two ranks
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, VGX_ROOT=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    assert "rank-0-ok" in p.stdout and "rank-1-ok" in p.stdout, p.stdout[-3000:]


def test_partition_properties():
    from voxgraph_b200 import api, build
    build.build()
    rs = np.random.RandomState(0)
    for n_ranks in (1, 2, 4, 8):
        counts = rs.randint(1, 20000, 300)
        owner = api.shard_constraints(n_ranks, counts)
        loads = np.bincount(owner, weights=counts, minlength=n_ranks)
        assert owner.min() >= 0 and owner.max() < n_ranks
        assert loads.max() - loads.min() <= counts.max()
        assert np.array_equal(owner, api.shard_constraints(n_ranks, counts))   # deterministic
    assert len(api.shard_constraints(4, [])) == 0
    # locality: with keys, each rank's constraints span a contiguous key range
    keys = rs.randint(0, 50, 300).astype(np.uint32)
    counts = np.full(300, 1000)
    owner = api.shard_constraints(4, counts, keys)
    spans = [(keys[owner == r].min(), keys[owner == r].max()) for r in range(4)]
    for r in range(3):
        assert spans[r][1] <= spans[r + 1][0]
    loads = np.bincount(owner, weights=counts, minlength=4)
    assert loads.max() - loads.min() <= 1000
