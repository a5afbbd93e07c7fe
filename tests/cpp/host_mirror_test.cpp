// Exercises the C++ host mirror (voxgraph_b200/host/voxgraph_b200.hpp) against the shared
// library on a GPU: a plane-field submap pair, Evaluate() through the cost-function mirror,
// a PoseGraph solve, and one integrated scan.  Prints "HOST_MIRROR_OK" on success.
#include <cmath>
#include <cstdio>
#include <vector>

#include "voxgraph_b200.hpp"

using namespace voxgraph_b200;

int main() {
  try {
    Context ctx(0);
    const float vs = 0.2f;
    const int vps = 16, vpb = vps * vps * vps;
    // layer sampling d(p) = z - 0.35 over blocks [-2,1]^2 x [-1,0]
    std::vector<int32_t> idx; std::vector<float> dist, w;
    for (int bz = -1; bz <= 0; ++bz) for (int by = -2; by <= 1; ++by) for (int bx = -2; bx <= 1; ++bx) {
      idx.insert(idx.end(), {bx, by, bz});
      for (int z = 0; z < vps; ++z) for (int y = 0; y < vps; ++y) for (int x = 0; x < vps; ++x) {
        dist.push_back((bz * vps + z + 0.5f) * vs - 0.35f);
        w.push_back(1.0f);
      }
    }
    const int nb = (int)idx.size() / 3;
    (void)vpb;
    ctx.uploadSubmap(0, vs, vps, nb, idx.data(), dist.data(), w.data());
    ctx.uploadSubmap(1, vs, vps, nb, idx.data(), dist.data(), w.data());
    // points on the plane z = 0.35 (distance 0), grid in x,y
    std::vector<float> xyz, pd, pw;
    for (int i = 0; i < 40; ++i) for (int j = 0; j < 40; ++j) {
      xyz.insert(xyz.end(), {-3.0f + 0.15f * i, -3.0f + 0.15f * j, 0.35f});
      pd.push_back(0.f); pw.push_back(1.f);
    }
    const int K = (int)pd.size();
    ctx.uploadRegistrationPoints(0, VGX_POINTS_ISOSURFACE, K, xyz.data(), pd.data(), pw.data());
    ctx.uploadRegistrationPoints(1, VGX_POINTS_ISOSURFACE, K, xyz.data(), pd.data(), pw.data());

    RegistrationCostFunction cost(ctx, 0, 1, RegistrationCostFunction::Config());
    if (cost.num_residuals() != K) { std::printf("bad num_residuals\n"); return 1; }
    double ref[4] = {0, 0, 0, 0}, read[4] = {0.1, -0.05, 0.2, 0.02};
    const double* params[2] = {ref, read};
    std::vector<double> r(K), j0(4 * K), j1(4 * K);
    double* jac[2] = {j0.data(), j1.data()};
    if (!cost.Evaluate(params, r.data(), jac)) { std::printf("Evaluate false\n"); return 1; }
    // reading moved up by 0.2: interpolated distance = (0.35 - 0.2) - 0.35 = -0.2 -> residual +0.2
    for (int i = 0; i < K; ++i)
      if (std::fabs(r[i] - 0.2) > 1e-5 || std::fabs(j0[4 * i + 2] + 1.0) > 1e-4 || std::fabs(j1[4 * i + 2] - 1.0) > 1e-4) {
        std::printf("residual/jacobian mismatch at %d: %g %g %g\n", i, r[i], j0[4 * i + 2], j1[4 * i + 2]);
        return 1;
      }

    PoseGraph graph(ctx);
    SubmapNodeConfig n0; n0.submap_id = 0; n0.set_constant = true;
    SubmapNodeConfig n1; n1.submap_id = 1; n1.T_mission_node_initial = {0.1, -0.05, 0.2, 0.02};
    graph.addSubmapNode(n0); graph.addSubmapNode(n1);
    RelativePoseConstraintConfig odo; odo.origin_submap_id = 0; odo.destination_submap_id = 1;
    odo.T_origin_destination = {0.1, -0.05, 0.0, 0.02};
    graph.addRelativePoseConstraint(odo);
    RegistrationConstraintConfig reg; reg.first_submap_id = 0; reg.second_submap_id = 1;
    graph.addRegistrationConstraint(reg);
    bool threw = false;
    try { RegistrationConstraintConfig bad; bad.first_submap_id = 0; bad.second_submap_id = 0; graph.addRegistrationConstraint(bad); }
    catch (const std::invalid_argument&) { threw = true; }
    if (!threw) { std::printf("self-constraint not rejected\n"); return 1; }
    // height measurement through a reference-frame node (information only on z, LDLT sqrt)
    ReferenceFrameNodeConfig frame; frame.reference_frame_id = 0;
    graph.addReferenceFrameNode(frame);
    AbsolutePoseConstraintConfig height; height.reference_frame_id = 0; height.submap_id = 1;
    height.information_matrix = InformationMatrix{}; height.information_matrix[10] = 2500.0;
    height.allow_semi_definite_information_matrix = true;
    height.T_ref_submap = {0.0, 0.0, 0.0, 0.0};
    graph.addAbsolutePoseConstraint(height);
    const std::array<double, 16> Sh = SqrtInformation(height.information_matrix, true);
    if (std::fabs(Sh[10] - 50.0) > 1e-12 || std::fabs(Sh[0]) + std::fabs(Sh[5]) + std::fabs(Sh[15]) > 0) {
      std::printf("LDLT sqrt information wrong\n"); return 1;
    }
    graph.solverOptions().parameter_tolerance = 1e-10;
    graph.optimize();
    const Pose4 p1 = graph.getSubmapPoses().at(1);
    const SolverSummary s = graph.getSolverSummaries().back();
    // the plane only constrains z; odometry keeps x, y, yaw and pulls z to 0 as well
    if (std::fabs(p1[2]) > 1e-4 || std::fabs(p1[0] - 0.1) > 1e-4 || !(s.final_cost < 1e-6 * s.initial_cost + 1e-12)) {
      std::printf("solve mismatch: z=%g x=%g cost %g -> %g\n", p1[2], p1[0], s.initial_cost, s.final_cost);
      return 1;
    }

    PointcloudIntegrator integ(ctx);
    integ.createSubmap(7, vs, vps, 1024);
    std::vector<float> pts;
    for (int i = 0; i < 2000; ++i) {
      const float a = 0.00314f * i;
      pts.insert(pts.end(), {6.0f * std::cos(a), 6.0f * std::sin(a), 0.3f * std::sin(3 * a)});
    }
    const float T[7] = {1, 0, 0, 0, 0, 0, 0};
    const vgx_tsdf_stats st = integ.integratePointcloud(7, T, 2000, pts.data());
    integ.finishSubmap(7);
    if (st.rays_valid != 2000 || st.voxel_updates <= 0 || st.blocks_allocated <= 0) { std::printf("integrate failed\n"); return 1; }
    std::printf("HOST_MIRROR_OK residuals=%d lm_iterations=%d tsdf_updates=%lld\n", K, s.iterations,
                (long long)st.voxel_updates);
    return 0;
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 2;
  }
}
