"""VGX_CHOL_DEBUG=1 python scripts/chol_probe.py : phase cycle counts of the shared-memory Cholesky on the
configs[1] pose graph (n = 196), printed by the library after each solve."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

sys.stdout = bench._REAL_STDOUT
w = dict(bench.WORKLOADS["config2"]); w["name"] = "config2"
sc = bench.build_scene(w)
from voxgraph_b200 import api  # noqa: E402
import time
ctx = api.Context(0)
P = bench.Problem(ctx, api, sc)
for rep in range(3):
    ctx.graph_set_poses(P.pinit)
    t0 = time.time()
    x, s = ctx.graph_solve(P.n_nodes, ctx.solver_options())
    print("solve %.3f ms, %d iterations" % ((time.time() - t0) * 1e3, s.iterations))
ctx.close()
