// Registration kernels: emit mode (Ceres layout, boundary b1) and fused reduce mode
// (per-constraint normal-equation sums, boundary b2).  Compiled with -fmad=false.
#include "registration.cuh"
#include "registration_kernels.h"

// ------------------------------------------------------------------ emit mode
// One thread per registration point; writes the normalised residual and the two 1x4
// Jacobian rows exactly as Evaluate leaves them for Ceres (cpp:254-291).
__global__ void __launch_bounds__(256)
reg_emit_kernel(RegConstraintDev C, RegPoseConst P, double* __restrict__ residuals,
                double* __restrict__ jac_ref, double* __restrict__ jac_read) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C.n) return;
  const float xi = __ldg(C.px + i), yi = __ldg(C.py + i), zi = __ldg(C.pz + i);
  const float dist = __ldg(C.pd + i), w = __ldg(C.pw + i);
  RegPointResult R;
  if (jac_ref || jac_read) R = vgx_reg_point<true>(C, P, xi, yi, zi, dist, w);
  else R = vgx_reg_point<false>(C, P, xi, yi, zi, dist, w);
  residuals[i] = R.r * C.factor;
  if (jac_ref) {
    double4 v;
    v.x = (double)R.jr[0] * C.factor; v.y = (double)R.jr[1] * C.factor;
    v.z = (double)R.jr[2] * C.factor; v.w = (double)R.jr[3] * C.factor;
    reinterpret_cast<double4*>(jac_ref)[i] = v;
  }
  if (jac_read) {
    double4 v;
    v.x = (double)(-R.jr[0]) * C.factor; v.y = (double)(-R.jr[1]) * C.factor;
    v.z = (double)(-R.jr[2]) * C.factor; v.w = (double)R.je3 * C.factor;
    reinterpret_cast<double4*>(jac_read)[i] = v;
  }
}

// ------------------------------------------------------------------ reduce mode
// Sums per tile, in double (products of float-valued doubles are exact, so fma == mul+add):
//   S[15] = upper triangle of sum j j^T over j = (jr0, jr1, jr2, jr3, je3)
//   g[5]  = sum j * r,   c = sum r^2
// The 8x8 block of the residual block follows from je[0..2] == -jr[0..2].

// Warp transpose-reduce: 32 per-lane values -> lane l holds the warp total of value l.
__device__ __forceinline__ double warp_transpose_reduce(double (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const double send = hi ? v[k] : v[k + off];
      const double keep = hi ? v[k + off] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <bool kJacobian>
__global__ void __launch_bounds__(VGX_REG_THREADS)
reg_reduce_kernel(const RegConstraintDev* __restrict__ constraints,
                  const RegPoseConst* __restrict__ poses, const RegTile* __restrict__ tiles,
                  double* __restrict__ partials) {
  __shared__ double s_part[VGX_REG_THREADS / 32][VGX_REG_NSUM];
  const RegTile T = tiles[blockIdx.x];
  const RegConstraintDev C = constraints[T.constraint];
  const RegPoseConst P = poses[T.constraint];
  double acc[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) acc[k] = 0.0;

  const int end = T.start + T.count;
  for (int i = T.start + threadIdx.x; i < end; i += VGX_REG_THREADS) {
    const float xi = __ldg(C.px + i), yi = __ldg(C.py + i), zi = __ldg(C.pz + i);
    const float dist = __ldg(C.pd + i), w = __ldg(C.pw + i);
    const RegPointResult R = vgx_reg_point<kJacobian>(C, P, xi, yi, zi, dist, w);
    acc[20] = fma(R.r, R.r, acc[20]);
    if (kJacobian) {
      const double j0 = (double)R.jr[0], j1 = (double)R.jr[1], j2 = (double)R.jr[2],
                   j3 = (double)R.jr[3], j4 = (double)R.je3;
      acc[0] = fma(j0, j0, acc[0]); acc[1] = fma(j0, j1, acc[1]); acc[2] = fma(j0, j2, acc[2]);
      acc[3] = fma(j0, j3, acc[3]); acc[4] = fma(j0, j4, acc[4]);
      acc[5] = fma(j1, j1, acc[5]); acc[6] = fma(j1, j2, acc[6]); acc[7] = fma(j1, j3, acc[7]);
      acc[8] = fma(j1, j4, acc[8]);
      acc[9] = fma(j2, j2, acc[9]); acc[10] = fma(j2, j3, acc[10]); acc[11] = fma(j2, j4, acc[11]);
      acc[12] = fma(j3, j3, acc[12]); acc[13] = fma(j3, j4, acc[13]);
      acc[14] = fma(j4, j4, acc[14]);
      acc[15] = fma(j0, R.r, acc[15]); acc[16] = fma(j1, R.r, acc[16]);
      acc[17] = fma(j2, R.r, acc[17]); acc[18] = fma(j3, R.r, acc[18]);
      acc[19] = fma(j4, R.r, acc[19]);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double tot = warp_transpose_reduce(acc, lane);
  if (lane < VGX_REG_NSUM) s_part[warp][lane] = tot;
  __syncthreads();
  if (threadIdx.x < VGX_REG_NSUM) {
    double s = 0;
#pragma unroll
    for (int wv = 0; wv < VGX_REG_THREADS / 32; ++wv) s += s_part[wv][threadIdx.x];
    partials[(size_t)blockIdx.x * VGX_REG_NSTRIDE + threadIdx.x] = s;
  }
}

// Fixed-order sum of the tile partials of each constraint, scaled by factor^2 (cpp:274-291).
__global__ void reg_finalize_kernel(const RegConstraintDev* __restrict__ constraints,
                                    const int* __restrict__ tile_begin,
                                    const double* __restrict__ partials,
                                    double* __restrict__ csum, int n_constraints) {
  const int c = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= n_constraints || lane >= VGX_REG_NSUM) return;
  const int t0 = tile_begin[c], t1 = tile_begin[c + 1];
  double s = 0;
  for (int t = t0; t < t1; ++t) s += partials[(size_t)t * VGX_REG_NSTRIDE + lane];
  const double f = constraints[c].factor;
  csum[(size_t)c * VGX_REG_NSTRIDE + lane] = s * (f * f);
}

__global__ void reg_pose_setup_kernel(const RegConstraintDev* __restrict__ constraints,
                                      const double* __restrict__ x, RegPoseConst* __restrict__ poses,
                                      int n_constraints) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_constraints) return;
  RegPoseConst P;
  vgx_reg_pose_setup<TrigDevice>(x + 4 * constraints[c].ref_node, x + 4 * constraints[c].read_node, P);
  poses[c] = P;
}

// ------------------------------------------------------------------ launch helpers
void vgx_launch_reg_pose_setup(cudaStream_t st, const RegConstraintDev* cons, const double* x,
                               RegPoseConst* poses, int n) {
  if (n <= 0) return;
  reg_pose_setup_kernel<<<(n + 127) / 128, 128, 0, st>>>(cons, x, poses, n);
}

void vgx_launch_reg_reduce(cudaStream_t st, const RegConstraintDev* cons, const RegPoseConst* poses,
                           const RegTile* tiles, int n_tiles, double* partials, bool jacobian) {
  if (n_tiles <= 0) return;
  if (jacobian)
    reg_reduce_kernel<true><<<n_tiles, VGX_REG_THREADS, 0, st>>>(cons, poses, tiles, partials);
  else
    reg_reduce_kernel<false><<<n_tiles, VGX_REG_THREADS, 0, st>>>(cons, poses, tiles, partials);
}

void vgx_launch_reg_finalize(cudaStream_t st, const RegConstraintDev* cons, const int* tile_begin,
                             const double* partials, double* csum, int n) {
  if (n <= 0) return;
  const int per_block = 4;
  reg_finalize_kernel<<<(n + per_block - 1) / per_block, 32 * per_block, 0, st>>>(cons, tile_begin,
                                                                                   partials, csum, n);
}

int vgx_fill_constraint(vgx_ctx* c, uint32_t ref_id, uint32_t read_id, const vgx_reg_config* cfg,
                        RegConstraintDev* out) {
  if (!cfg) VGX_FAIL(c, VGX_ERR_INVALID, "null registration config");
  if (cfg->sampling_ratio != -1.0f)
    VGX_FAIL(c, VGX_ERR_INVALID,
             "sampling_ratio != -1 (random re-sampling per Evaluate) is not supported; "
             "pre-sample the registration points on the host");
  if (cfg->registration_point_type < 0 || cfg->registration_point_type > 1)
    VGX_FAIL(c, VGX_ERR_INVALID, "invalid registration_point_type");
  if (ref_id == read_id) VGX_FAIL(c, VGX_ERR_INVALID, "cannot constrain a submap to itself");
  VgxSubmap* ref = c->find(ref_id);
  VgxSubmap* rd = c->find(read_id);
  if (!ref || !rd) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "registration constraint: unknown submap");
  if (!rd->finished || !rd->d_view)
    VGX_FAIL(c, VGX_ERR_INVALID, "registration constraint: reading submap is not finished");
  const VgxPoints& p = ref->points[cfg->registration_point_type];
  RegConstraintDev C;
  C.px = p.x; C.py = p.y; C.pz = p.z; C.pd = p.dist; C.pw = p.w;
  C.n = p.n;
  C.ref_node = -1; C.read_node = -1;
  C.hash = rd->hash;
  C.view = rd->d_view;
  C.voxel_size = rd->voxel_size; C.voxel_size_inv = rd->voxel_size_inv;
  C.block_size = rd->block_size; C.block_size_inv = rd->block_size_inv;
  C.vps = rd->vps;
  C.vps_shift = 0;
  while ((1 << C.vps_shift) < rd->vps) C.vps_shift++;
  C.factor = (p.sum_w != 0.0) ? (double)p.n / p.sum_w : 0.0;
  C.no_corr = cfg->no_correspondence_cost;
  *out = C;
  return VGX_OK;
}

// ------------------------------------------------------------------ C-ABI: b1
extern "C" void vgx_reg_config_default(vgx_reg_config* cfg) {
  if (!cfg) return;
  cfg->registration_point_type = VGX_POINTS_ISOSURFACE;  // h:20-21
  cfg->no_correspondence_cost = 0.0;                      // h:32
  cfg->sampling_ratio = -1.0f;                            // h:28
}

extern "C" int vgx_reg_num_residuals(vgx_ctx* c, uint32_t ref_id, const vgx_reg_config* cfg, int* n) {
  if (!c || !cfg || !n) return VGX_ERR_INVALID;
  VgxSubmap* ref = c->find(ref_id);
  if (!ref) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_reg_num_residuals: unknown submap");
  if (cfg->registration_point_type < 0 || cfg->registration_point_type > 1) return VGX_ERR_INVALID;
  *n = ref->points[cfg->registration_point_type].n;
  return VGX_OK;
}

extern "C" int vgx_reg_eval_emit(vgx_ctx* c, uint32_t ref_id, uint32_t read_id,
                                 const vgx_reg_config* cfg, const double ref_pose[4],
                                 const double read_pose[4], double* residuals, double* jac_ref,
                                 double* jac_read) {
  if (!c || !ref_pose || !read_pose || !residuals) return VGX_ERR_INVALID;
  VGX_CUDA(c, cudaSetDevice(c->device));
  RegConstraintDev C;
  int rc = vgx_fill_constraint(c, ref_id, read_id, cfg, &C);
  if (rc != VGX_OK) return rc;
  if (C.n == 0) return VGX_ZERO_WEIGHT;  // summed weight 0 -> Evaluate returns false
  if (C.factor == 0.0) return VGX_ZERO_WEIGHT;
  RegPoseConst P;
  vgx_reg_pose_setup<TrigHostLibm>(ref_pose, read_pose, P);
  const size_t K = (size_t)C.n;
  const size_t Kp = (K + 3) & ~(size_t)3;  // keep the Jacobian blocks 32-byte aligned
  rc = c->ensure_scratch(9 * Kp * sizeof(double));
  if (rc != VGX_OK) return rc;
  double* d_r = (double*)c->d_scratch;
  double* d_jr = jac_ref ? d_r + Kp : nullptr;
  double* d_je = jac_read ? d_r + 5 * Kp : nullptr;
  {
    VgxLaunchScope scope(c, 1);
    reg_emit_kernel<<<(unsigned)((K + 255) / 256), 256, 0, c->stream>>>(C, P, d_r, d_jr, d_je);
  }
  VGX_CUDA(c, cudaGetLastError());
  VGX_CUDA(c, cudaMemcpyAsync(residuals, d_r, K * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (jac_ref)
    VGX_CUDA(c, cudaMemcpyAsync(jac_ref, d_jr, 4 * K * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (jac_read)
    VGX_CUDA(c, cudaMemcpyAsync(jac_read, d_je, 4 * K * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  VGX_CUDA(c, cudaStreamSynchronize(c->stream));
  return VGX_OK;
}
