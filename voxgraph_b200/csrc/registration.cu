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

// The sums are the 6x6 Gram matrix of v = (jr0, jr1, jr2, jr3, je3, r) over the points, formed
// on the FP64 tensor cores: each warp stages its 32 points' vectors in shared memory and issues
// eight mma.sync.m8n8k4.f64 (A = B^T = 4 points x 8 components), so the reduction over points
// happens inside the MMA and each lane keeps just two accumulators.
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

#define VGX_STAGE_STRIDE 36  // doubles per component row (32 points + pad: 2-way = optimal for 8 B)

// Persistent CTAs: the residual index space is cut evenly over the grid (148 x resident CTAs per
// SM); a CTA walks its tiles (a tile never straddles two residual blocks).  Per tile: the
// reading submap's dense block grid is staged in shared memory, points stream through
// transform -> voxel index -> grid lookup -> one octet -> residual/Jacobian -> Gram MMA.
template <bool kJacobian>
__global__ void __launch_bounds__(VGX_REG_THREADS, VGX_REG_MIN_BLOCKS)
reg_reduce_kernel(const RegConstraintDev* __restrict__ constraints,
                  const RegPoseConst* __restrict__ poses, const RegTile* __restrict__ tiles,
                  const int* __restrict__ cta_tile_begin, const int* __restrict__ tile_begin,
                  int* __restrict__ counters, double* __restrict__ partials,
                  double* __restrict__ csum, int grid_capacity) {
  constexpr int kWarps = VGX_REG_THREADS / 32;
  __shared__ double s_stage[kWarps][6][VGX_STAGE_STRIDE];
  __shared__ double s_gram[kWarps][64];
  __shared__ int s_last;
  extern __shared__ int32_t s_grid[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane >> 2, tig = lane & 3;
  int grid_of = -1;  // constraint whose block grid is currently staged

  for (int tile = cta_tile_begin[blockIdx.x]; tile < cta_tile_begin[blockIdx.x + 1]; ++tile) {
    const RegTile T = tiles[tile];
    const RegConstraintDev C = constraints[T.constraint];
    const RegPoseConst P = poses[T.constraint];
    const int cells = C.gd0 * C.gd1 * C.gd2;
    const bool use_grid = C.grid != nullptr && cells <= grid_capacity;
    if (use_grid && grid_of != T.constraint) {
      __syncthreads();  // previous tile's readers are done
      for (int k = threadIdx.x; k < cells; k += VGX_REG_THREADS) s_grid[k] = __ldg(C.grid + k);
      grid_of = T.constraint;
      __syncthreads();
    }
    double d0 = 0.0, d1 = 0.0;
    const int end = T.start + T.count;
    const size_t vox_shift = 3 * C.vps_shift;
#if VGX_REG_PREFETCH
    // software pipeline: the next round's point is loaded while this round is computed
    float nx, ny, nz, nd, nw;
    {
      const int i0 = T.start + threadIdx.x;
      const int ic0 = i0 < end ? i0 : T.start;
      nx = __ldg(C.px + ic0); ny = __ldg(C.py + ic0); nz = __ldg(C.pz + ic0);
      nd = __ldg(C.pd + ic0); nw = __ldg(C.pw + ic0);
    }
#endif
    for (int base = T.start; base < end; base += VGX_REG_THREADS) {
      const int i = base + threadIdx.x;
      const bool act = i < end;
#if VGX_REG_PREFETCH
      const float xi = nx, yi = ny, zi = nz, dist = nd, w = nw;
      {
        const int in = i + VGX_REG_THREADS;
        const int icn = in < end ? in : T.start;
        nx = __ldg(C.px + icn); ny = __ldg(C.py + icn); nz = __ldg(C.pz + icn);
        nd = __ldg(C.pd + icn); nw = __ldg(C.pw + icn);
      }
#else
      const int ic = act ? i : T.start;
      const float xi = __ldg(C.px + ic), yi = __ldg(C.py + ic), zi = __ldg(C.pz + ic);
      const float dist = __ldg(C.pd + ic), w = __ldg(C.pw + ic);
#endif
      float p0, p1, p2;
      vgx_reg_transform(P, xi, yi, zi, p0, p1, p2);
      RegLocate L;
      int slot;
      if (use_grid) {
        vgx_locate<true>(C, p0, p1, p2, L, s_grid);
        slot = L.slot;
      } else {
        vgx_locate<false>(C, p0, p1, p2, L);
        slot = vgx_resolve(C, L);
      }
      const bool found = slot >= 0;
      const size_t lin = ((size_t)(found ? slot : 0) << vox_shift) + L.lin;
      const float4* o = reinterpret_cast<const float4*>(C.view) + 2 * lin;
#if VGX_REG_STREAM_OCTETS
      // an octet is touched once per evaluation: keep it out of L1 (evict-first)
      const float4 lo = __ldcs(o), hi = __ldcs(o + 1);
#else
      const float4 lo = __ldg(o), hi = __ldg(o + 1);
#endif
      const float d[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      const bool ok = found && vgx_octet_ok(d);
      const RegPointResult R = vgx_reg_math<kJacobian>(C, P, xi, yi, dist, w, ok, d, L.ox, L.oy, L.oz);
      double v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0;
      if (act) {
        v5 = R.r;
        if (kJacobian) {
          v0 = (double)R.jr[0]; v1 = (double)R.jr[1]; v2 = (double)R.jr[2];
          v3 = (double)R.jr[3]; v4 = (double)R.je3;
        }
      }
      if (kJacobian) {
        s_stage[warp][0][lane] = v0; s_stage[warp][1][lane] = v1; s_stage[warp][2][lane] = v2;
        s_stage[warp][3][lane] = v3; s_stage[warp][4][lane] = v4; s_stage[warp][5][lane] = v5;
        __syncwarp();
#pragma unroll
        for (int t4 = 0; t4 < 8; ++t4) {
          const double a = (grp < 6) ? s_stage[warp][grp][4 * t4 + tig] : 0.0;
          dmma_8x8x4(d0, d1, a, a);
        }
        __syncwarp();
      } else {
        d0 = fma(v5, v5, d0);
      }
    }
    // ---- tile epilogue: warp Gram fragments -> 21 sums -> partials[tile]
    if (kJacobian) {
      s_gram[warp][grp * 8 + 2 * tig] = d0;
      s_gram[warp][grp * 8 + 2 * tig + 1] = d1;
    } else {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) d0 += __shfl_xor_sync(0xffffffffu, d0, off);
      if (lane == 0) s_gram[warp][0] = d0;
    }
    __syncthreads();
    if (threadIdx.x < VGX_REG_NSUM) {
      // entry e of the 21 sums -> (row, col) of the Gram matrix
      int row = 5, col = 5;
      const int e = threadIdx.x;
      if (e < 15) {
        int p = 0, rem = e;
        while (rem >= 5 - p) { rem -= 5 - p; ++p; }
        row = p; col = p + rem;
      } else if (e < 20) {
        row = e - 15; col = 5;
      }
      double s = 0;
      if (kJacobian) {
#pragma unroll
        for (int wv = 0; wv < kWarps; ++wv) s += s_gram[wv][row * 8 + col];
      } else if (e == 20) {
#pragma unroll
        for (int wv = 0; wv < kWarps; ++wv) s += s_gram[wv][0];
      }
      partials[(size_t)tile * VGX_REG_NSTRIDE + e] = s;
    }
    // The last tile of a constraint to finish sums the constraint's partials in tile order
    // (bit-reproducible) and applies factor^2 (cpp:274-291).
    const int t0 = tile_begin[T.constraint], t1 = tile_begin[T.constraint + 1];
    if (t1 - t0 == 1) {
      if (threadIdx.x < VGX_REG_NSUM)  // single tile: no ticket needed (same thread wrote the partial)
        csum[(size_t)T.constraint * VGX_REG_NSTRIDE + threadIdx.x] =
            partials[(size_t)tile * VGX_REG_NSTRIDE + threadIdx.x] * (C.factor * C.factor);
      __syncthreads();
    } else {
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        const int done = atomicAdd(counters + T.constraint, 1);
        s_last = (done == t1 - t0 - 1);
      }
      __syncthreads();
      if (s_last) {
        __threadfence();
        if (threadIdx.x < VGX_REG_NSUM) {
          double s = 0;
          for (int t = t0; t < t1; ++t) s += __ldcg(partials + (size_t)t * VGX_REG_NSTRIDE + threadIdx.x);
          csum[(size_t)T.constraint * VGX_REG_NSTRIDE + threadIdx.x] = s * (C.factor * C.factor);
        }
        if (threadIdx.x == 0) counters[T.constraint] = 0;
      }
    }
  }
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
                           const RegTile* tiles, int n_ctas, const int* cta_tile_begin,
                           const int* tile_begin, int* counters, double* partials, double* csum,
                           int grid_capacity, bool jacobian) {
  if (n_ctas <= 0) return;
  const size_t smem = sizeof(int32_t) * (size_t)grid_capacity;
  if (jacobian)
    reg_reduce_kernel<true><<<n_ctas, VGX_REG_THREADS, smem, st>>>(cons, poses, tiles, cta_tile_begin,
                                                                  tile_begin, counters, partials, csum,
                                                                  grid_capacity);
  else
    reg_reduce_kernel<false><<<n_ctas, VGX_REG_THREADS, smem, st>>>(cons, poses, tiles, cta_tile_begin,
                                                                   tile_begin, counters, partials, csum,
                                                                   grid_capacity);
}

int vgx_reg_resident_ctas(int device) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return sms * VGX_REG_MIN_BLOCKS;
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
  C.grid = rd->d_grid;
  C.gmin0 = rd->grid_min[0]; C.gmin1 = rd->grid_min[1]; C.gmin2 = rd->grid_min[2];
  C.gd0 = rd->grid_dim[0]; C.gd1 = rd->grid_dim[1]; C.gd2 = rd->grid_dim[2];
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
