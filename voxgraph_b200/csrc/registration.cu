// Registration kernels: emit mode (Ceres layout, boundary b1) and fused reduce mode
// (per-constraint normal-equation sums, boundary b2).  Compiled with -fmad=false.
#include "registration.cuh"
#include "registration_kernels.h"

#include <stdlib.h>

// ------------------------------------------------------------------ emit mode
// One thread per registration point; writes the normalised residual and the two 1x4
// Jacobian rows exactly as Evaluate leaves them for Ceres (cpp:254-291).
__global__ void __launch_bounds__(256)
reg_emit_kernel(RegConstraintDev C, RegPoseConst P, double* __restrict__ residuals,
                double* __restrict__ jac_ref, double* __restrict__ jac_read) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C.n) return;
  const float* pp = C.pts + vgx_pt_index((size_t)i, 0);
  const float xi = __ldg(pp), yi = __ldg(pp + 32), zi = __ldg(pp + 64);
  const float dist = __ldg(pp + 96), w = __ldg(pp + 128);
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

// ---- TMA (1-D bulk copy) + mbarrier primitives, sm_100a PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "VGX_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra VGX_DONE;\n"
      "bra VGX_WAIT;\n"
      "VGX_DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk copy completing on an mbarrier (SASS UBLKCP); 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Registration reduce kernel, v6: warp-level dynamic scheduling.
//
// The residual index space of every constraint is cut into tiles of VGX_REG_TILE_UNITS 32-point
// units (128 points).  The cost of a tile varies by 4x with the share of its points that land in
// the reading submap (overlapping submaps share only part of their surface and consecutive
// points are spatially coherent), so tiles are NOT pre-assigned: every warp of the persistent
// grid draws tile tickets from a global counter and runs its own software pipeline, two tiles
// deep, with no block-level synchronisation at all:
//   * ticket t+2 is drawn and its points (TILE_UNITS x 640 contiguous bytes of the unit-major
//     AoSoA layout) are requested with ONE cp.async.bulk (TMA 1-D) into the warp's 2-slot
//     shared-memory ring, completing on an mbarrier, while tile t is processed;
//   * per tile the constraint's descriptor + float pose block are staged in the warp's own
//     shared-memory copy (lanes load one word each);
//   * stage A(u+1): transform -> containing block -> block grid (global memory, L1/L2 resident,
//     <= 16 KB per submap) -> base corner voxel -> the octet's two LDG.128 are ISSUED;
//     stage B(u): the octet issued one unit earlier is consumed: B1 coefficients, residual,
//     Jacobian, Gram staging + 8 DMMA.  Two register sets alternate so the in-flight octet is
//     never copied.  A warp whose 32 points all miss the reading submap (getVoxelsAndQVector
//     fails on its first block lookup) skips the unit after ~60 instructions;
//   * the tile's 21 sums go to partials[tile]; the warp that finishes a constraint's last tile
//     (atomic ticket) adds the constraint's partials in tile order -> bit-reproducible results
//     whatever warp processed which tile.
// The last unit of a constraint is zero-padded to 32 points (weight 0 -> zero contribution).
#define VGX_REG_UNIT_FLOATS 160
#define VGX_REG_TILE_BYTES (VGX_REG_TILE_UNITS * 640)

struct RegPipeRegs {   // what stage A hands to stage B
  float4 lo, hi;       // the octet (NaN when the block is missing)
  float ox, oy, oz;
  bool found;
};

// block grid of the reading submap in global memory (int32 slots, -1 = no block)
__device__ __forceinline__ int vgx_grid_slot_g(const RegConstraintDev& C, int b0, int b1, int b2) {
  const int g0 = b0 - C.gmin0, g1 = b1 - C.gmin1, g2 = b2 - C.gmin2;
  const bool in = (unsigned)g0 < (unsigned)C.gd0 && (unsigned)g1 < (unsigned)C.gd1 &&
                  (unsigned)g2 < (unsigned)C.gd2;
  return in ? __ldg(C.grid + (g2 * C.gd1 + g1) * C.gd0 + g0) : -1;
}

template <bool kJacobian>
__global__ void __launch_bounds__(VGX_REG_THREADS, VGX_REG_MIN_BLOCKS)
reg_reduce_kernel(const RegConstraintDev* __restrict__ constraints,
                  const RegPoseConst* __restrict__ poses, const RegTile* __restrict__ tiles,
                  int n_tiles, const int* __restrict__ tile_begin, int* __restrict__ counters,
                  int* __restrict__ sched /* [0] next ticket, [1] warps that have left */,
                  double* __restrict__ partials, double* __restrict__ csum) {
  constexpr int kWarps = VGX_REG_THREADS / 32;
  constexpr int kTU = VGX_REG_TILE_UNITS;
  __shared__ __align__(128) float s_ring[kWarps][2][kTU * VGX_REG_UNIT_FLOATS];
  __shared__ __align__(8) unsigned long long s_bar[kWarps][2];
  __shared__ double s_stage[kWarps][6][VGX_STAGE_STRIDE];
  __shared__ double s_gram[kWarps][64];
  __shared__ __align__(16) RegConstraintDev s_Cw[kWarps];
  __shared__ __align__(16) RegPoseConst s_Pw[kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane >> 2, tig = lane & 3;
  const RegConstraintDev& C = s_Cw[warp];
  if (lane == 0) {
    mbar_init(smem_u32(&s_bar[warp][0]), 1);
    mbar_init(smem_u32(&s_bar[warp][1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint32_t ring0 = smem_u32(&s_ring[warp][0][0]);
  const uint32_t bar0 = smem_u32(&s_bar[warp][0]);
  const float* ringf = &s_ring[warp][0][0] + lane;
  double* stage_w = &s_stage[warp][0][lane];
  const double* stage_r = &s_stage[warp][grp < 6 ? grp : 0][tig];

  auto draw = [&]() -> int {  // next tile ticket, warp-uniform
    int t = 0;
    if (lane == 0) t = atomicAdd(sched, 1);
    return __shfl_sync(0xffffffffu, t, 0);
  };
  // request tile t's points into ring slot `slot` (lane 0); returns the tile record to all lanes
  auto request = [&](int t, uint32_t slot) -> RegTile {
    RegTile T;
    T.constraint = -1; T.start = 0; T.count = 0; T.pad = 0;
    if (t < n_tiles) {
      const int4 raw = __ldg(reinterpret_cast<const int4*>(tiles + t));
      T.constraint = raw.x; T.start = raw.y; T.count = raw.z;
      if (lane == 0) {
        const float* src = reinterpret_cast<const float*>(__ldg(reinterpret_cast<const unsigned long long*>(
                               &constraints[T.constraint].pts))) +
                           (size_t)(T.start >> 5) * VGX_REG_UNIT_FLOATS;
        const uint32_t bytes = 640u * (uint32_t)((T.count + 31) >> 5);
        mbar_expect_tx(bar0 + 8u * slot, bytes);
        tma_load_1d(ring0 + (uint32_t)VGX_REG_TILE_BYTES * slot, src, bytes, bar0 + 8u * slot);
      }
    }
    return T;
  };

  uint32_t seq = 0;  // tiles consumed by this warp: ring slot = seq & 1, barrier parity = (seq >> 1) & 1
  int t_cur = draw();
  RegTile T_cur = request(t_cur, 0);
  int t_nxt = draw();
  RegTile T_nxt = request(t_nxt, 1);

  while (t_cur < n_tiles) {
    const uint32_t slot = seq & 1u;
    // ---- stage the constraint descriptor + pose block in the warp's shared-memory copy
    {
      constexpr int kWc = (int)(sizeof(RegConstraintDev) / 4), kWp = (int)(sizeof(RegPoseConst) / 4);
      const uint32_t* gc = reinterpret_cast<const uint32_t*>(constraints + T_cur.constraint);
      const uint32_t* gp = reinterpret_cast<const uint32_t*>(poses + T_cur.constraint);
      uint32_t* sc = reinterpret_cast<uint32_t*>(&s_Cw[warp]);
      uint32_t* sp = reinterpret_cast<uint32_t*>(&s_Pw[warp]);
      for (int k = lane; k < kWc; k += 32) sc[k] = __ldg(gc + k);
      if (lane < kWp) sp[lane] = __ldg(gp + lane);
      __syncwarp();
    }
    const RegPoseConst P = s_Pw[warp];
    const int n_units = (T_cur.count + 31) >> 5;
    const bool use_grid = C.grid != nullptr;
    const size_t vox_shift = 3 * C.vps_shift;
    const float4* view = reinterpret_cast<const float4*>(C.view);
    const bool gram_always = C.no_corr != 0.0;  // r = w * no_correspondence_cost without a match
    const float* tile_ring = ringf + slot * (kTU * VGX_REG_UNIT_FLOATS);
    double d0 = 0.0, d1 = 0.0;
    mbar_wait(bar0 + 8u * slot, (seq >> 1) & 1u);  // this tile's points have landed

    // stage A: everything up to the ISSUE of the octet loads
    auto stage_a = [&](int u, RegPipeRegs& R) {
      const float* sp = tile_ring + u * VGX_REG_UNIT_FLOATS;
      float p0, p1, p2;
      vgx_reg_transform(P, sp[0], sp[32], sp[64], p0, p1, p2);
      const float qnan = __int_as_float(0x7fc00000);
      R.lo = make_float4(qnan, qnan, qnan, qnan);
      R.hi = R.lo;
      R.found = false;
      R.ox = R.oy = R.oz = 0.f;
      int b0, b1, b2;
      vgx_block_index(C, p0, p1, p2, b0, b1, b2);
      RegLocate L;
      int slot_b;
      if (use_grid) {
        // getVoxelsAndQVector fails right away when the block that contains pos does not exist
        const bool hit = vgx_grid_slot_g(C, b0, b1, b2) >= 0;
        if (!__any_sync(0xffffffffu, hit)) return;
        vgx_locate_in_block<false>(C, p0, p1, p2, b0, b1, b2, L, nullptr, true);
        slot_b = hit ? vgx_grid_slot_g(C, L.b0, L.b1, L.b2) : -1;
      } else {
        vgx_locate_in_block<false>(C, p0, p1, p2, b0, b1, b2, L);
        slot_b = vgx_resolve(C, L);
      }
      R.found = slot_b >= 0;
      R.ox = L.ox; R.oy = L.oy; R.oz = L.oz;
      if (R.found) {
        const float4* o = view + 2 * (((size_t)slot_b << vox_shift) + (size_t)L.lin);
#if VGX_REG_STREAM_OCTETS
        R.lo = __ldcs(o); R.hi = __ldcs(o + 1);
#else
        R.lo = __ldg(o); R.hi = __ldg(o + 1);
#endif
      }
    };
    // stage B: consume the octet issued one unit ago
    auto stage_b = [&](int u, const RegPipeRegs& R) {
      const float* sp = tile_ring + u * VGX_REG_UNIT_FLOATS;
      const float d[8] = {R.lo.x, R.lo.y, R.lo.z, R.lo.w, R.hi.x, R.hi.y, R.hi.z, R.hi.w};
      bool ok = false;
      if (__any_sync(0xffffffffu, R.found)) ok = R.found && vgx_octet_ok(d);
      if (gram_always || __any_sync(0xffffffffu, ok)) {
        const float xi = sp[0], yi = sp[32], dist = sp[96], w = sp[128];
        const RegPointResult Q = vgx_reg_math<kJacobian>(C, P, xi, yi, dist, w, ok, d, R.ox, R.oy, R.oz);
        if (kJacobian) {
          stage_w[0 * VGX_STAGE_STRIDE] = (double)Q.jr[0];
          stage_w[1 * VGX_STAGE_STRIDE] = (double)Q.jr[1];
          stage_w[2 * VGX_STAGE_STRIDE] = (double)Q.jr[2];
          stage_w[3 * VGX_STAGE_STRIDE] = (double)Q.jr[3];
          stage_w[4 * VGX_STAGE_STRIDE] = (double)Q.je3;
          stage_w[5 * VGX_STAGE_STRIDE] = Q.r;
          __syncwarp();
#pragma unroll
          for (int t4 = 0; t4 < 8; ++t4) {
            const double a = (grp < 6) ? stage_r[4 * t4] : 0.0;
            dmma_8x8x4(d0, d1, a, a);
          }
          __syncwarp();
        } else {
          d0 = fma(Q.r, Q.r, d0);
        }
      }
    };

    {
      RegPipeRegs Ra, Rb;
      stage_a(0, Ra);
      int u = 0;
      for (; u + 2 <= n_units; u += 2) {
        stage_a(u + 1, Rb);
        stage_b(u, Ra);
        if (u + 2 < n_units) stage_a(u + 2, Ra);
        stage_b(u + 1, Rb);
      }
      if (u < n_units) stage_b(u, Ra);
    }
    __syncwarp();  // every lane is done with the ring slot and the constants
    // ---- the slot is free: draw the ticket after next and request its points
    const int t_new = draw();
    const RegTile T_new = request(t_new, slot);

    // ---- tile epilogue: warp Gram fragment -> 21 sums -> partials[tile]
    if (kJacobian) {
      s_gram[warp][grp * 8 + 2 * tig] = d0;
      s_gram[warp][grp * 8 + 2 * tig + 1] = d1;
    } else {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) d0 += __shfl_xor_sync(0xffffffffu, d0, off);
    }
    __syncwarp();
    const int cidx = T_cur.constraint;
    const double factor = C.factor;
    if (lane < VGX_REG_NSUM) {
      // entry e of the 21 sums -> (row, col) of the Gram matrix
      int row = 5, col = 5;
      const int e = lane;
      if (e < 15) {
        int p = 0, rem = e;
        while (rem >= 5 - p) { rem -= 5 - p; ++p; }
        row = p; col = p + rem;
      } else if (e < 20) {
        row = e - 15; col = 5;
      }
      double sv = 0;
      if (kJacobian) sv = s_gram[warp][row * 8 + col];
      else if (e == 20) sv = d0;
      partials[(size_t)t_cur * VGX_REG_NSTRIDE + e] = sv;
      __threadfence();
    }
    __syncwarp();
    // the warp that completes the constraint adds its partials in tile order (cpp:274-291: factor^2)
    const int t0 = tile_begin[cidx], t1 = tile_begin[cidx + 1];
    int last = 0;
    if (lane == 0) last = (atomicAdd(counters + cidx, 1) == t1 - t0 - 1);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      __threadfence();
      if (lane < VGX_REG_NSUM) {
        double sv = 0;
        for (int t = t0; t < t1; ++t) sv += __ldcg(partials + (size_t)t * VGX_REG_NSTRIDE + lane);
        csum[(size_t)cidx * VGX_REG_NSTRIDE + lane] = sv * (factor * factor);
      }
      if (lane == 0) counters[cidx] = 0;
    }
    __syncwarp();
    t_cur = t_nxt; T_cur = T_nxt;
    t_nxt = t_new; T_nxt = T_new;
    ++seq;
  }
  // ---- the last warp to leave re-arms the scheduler for the next launch
  if (lane == 0) {
    const int total = (int)(gridDim.x * kWarps);
    __threadfence();
    if (atomicAdd(sched + 1, 1) == total - 1) {
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
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
                           const RegTile* tiles, int n_tiles, int n_ctas, const int* tile_begin,
                           int* counters, int* sched, double* partials, double* csum, bool jacobian) {
  if (n_ctas <= 0 || n_tiles <= 0) return;
  if (jacobian)
    reg_reduce_kernel<true><<<n_ctas, VGX_REG_THREADS, 0, st>>>(cons, poses, tiles, n_tiles, tile_begin,
                                                               counters, sched, partials, csum);
  else
    reg_reduce_kernel<false><<<n_ctas, VGX_REG_THREADS, 0, st>>>(cons, poses, tiles, n_tiles, tile_begin,
                                                                counters, sched, partials, csum);
}

int vgx_reg_resident_ctas(int device) {
  int sms = 148, per_sm = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reg_reduce_kernel<true>, VGX_REG_THREADS,
                                                    0) != cudaSuccess || per_sm <= 0)
    per_sm = VGX_REG_MIN_BLOCKS;
  static const char* env = getenv("VGX_REG_CTAS_PER_SM");  // tuning override
  if (env && atoi(env) > 0) per_sm = atoi(env);
  return sms * per_sm;
}

// cpp:45-55: num_residuals = int(sampling_ratio * size()) (float product) or size()
static int reg_num_residuals(const VgxPoints& p, const vgx_reg_config* cfg) {
  if (cfg->sampling_ratio != -1.0f) return (int)(cfg->sampling_ratio * (float)(size_t)p.n);
  return p.n;
}

__global__ void reg_gather_samples_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                          int n, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ((n + 31) & ~31)) return;
  float x = 0.f, y = 0.f, z = 0.f, d = 0.f, w = 0.f;
  if (i < n) {
    const float* sp = src + vgx_pt_index((size_t)idx[i], 0);
    x = sp[0]; y = sp[32]; z = sp[64]; d = sp[96];
    w = 1.0f;  // cpp:121 registration_point.weight = 1
  }
  float* dp = dst + vgx_pt_index((size_t)i, 0);
  dp[0] = x; dp[32] = y; dp[64] = z; dp[96] = d; dp[128] = w;
}

void vgx_launch_reg_gather_samples(cudaStream_t st, const float* src, const int32_t* d_idx, int n,
                                   float* dst) {
  if (n <= 0) return;
  const int padded = (n + 31) & ~31;
  reg_gather_samples_kernel<<<(padded + 255) / 256, 256, 0, st>>>(src, d_idx, n, dst);
}

int vgx_fill_constraint(vgx_ctx* c, uint32_t ref_id, uint32_t read_id, const vgx_reg_config* cfg,
                        RegConstraintDev* out, bool* sampled) {
  if (!cfg) VGX_FAIL(c, VGX_ERR_INVALID, "null registration config");
  if (cfg->registration_point_type < 0 || cfg->registration_point_type > 1)
    VGX_FAIL(c, VGX_ERR_INVALID, "invalid registration_point_type");
  if (cfg->sampling_ratio != -1.0f && !(cfg->sampling_ratio >= 0.0f))
    VGX_FAIL(c, VGX_ERR_INVALID, "sampling_ratio must be -1 (all points) or >= 0");
  if (ref_id == read_id) VGX_FAIL(c, VGX_ERR_INVALID, "cannot constrain a submap to itself");
  VgxSubmap* ref = c->find(ref_id);
  VgxSubmap* rd = c->find(read_id);
  if (!ref || !rd) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "registration constraint: unknown submap");
  if (!rd->finished || !rd->d_view)
    VGX_FAIL(c, VGX_ERR_INVALID, "registration constraint: reading submap is not finished");
  const VgxPoints& p = ref->points[cfg->registration_point_type];
  const bool is_sampled = cfg->sampling_ratio != -1.0f;
  RegConstraintDev C;
  C.pts = is_sampled ? nullptr : p.data;
  C.n = reg_num_residuals(p, cfg);
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
  // cpp:274 factor = num_residuals / summed_reference_weight; sampled points all weigh 1
  if (is_sampled) C.factor = (C.n > 0 && p.n > 0) ? 1.0 : 0.0;
  else C.factor = (p.sum_w != 0.0) ? (double)p.n / p.sum_w : 0.0;
  C.no_corr = cfg->no_correspondence_cost;
  *out = C;
  if (sampled) *sampled = is_sampled;
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
  if (cfg->sampling_ratio != -1.0f && !(cfg->sampling_ratio >= 0.0f)) return VGX_ERR_INVALID;
  *n = reg_num_residuals(ref->points[cfg->registration_point_type], cfg);
  return VGX_OK;
}

extern "C" int vgx_reg_eval_emit(vgx_ctx* c, uint32_t ref_id, uint32_t read_id,
                                 const vgx_reg_config* cfg, const double ref_pose[4],
                                 const double read_pose[4], double* residuals, double* jac_ref,
                                 double* jac_read) {
  if (!c || !ref_pose || !read_pose || !residuals) return VGX_ERR_INVALID;
  VGX_CUDA(c, cudaSetDevice(c->device));
  RegConstraintDev C;
  bool sampled = false;
  int rc = vgx_fill_constraint(c, ref_id, read_id, cfg, &C, &sampled);
  if (rc != VGX_OK) return rc;
  if (C.n == 0) return VGX_ZERO_WEIGHT;  // summed weight 0 -> Evaluate returns false
  if (C.factor == 0.0) return VGX_ZERO_WEIGHT;
  const size_t K = (size_t)C.n;
  const size_t Kp = (K + 3) & ~(size_t)3;  // keep the Jacobian blocks 32-byte aligned
  const size_t units = (K + 31) / 32;
  const size_t sample_bytes = sampled ? ((units * VGX_PT_UNIT_FLOATS * sizeof(float) + Kp * sizeof(int32_t) + 255) & ~(size_t)255) : 0;
  rc = c->ensure_scratch(9 * Kp * sizeof(double) + sample_bytes);
  if (rc != VGX_OK) return rc;
  if (sampled) {
    // cpp:118-122: every Evaluate draws its points anew from the reference submap's sampler
    VgxPoints& p = c->find(ref_id)->points[cfg->registration_point_type];
    rc = c->ensure_pinned(Kp * sizeof(int32_t));
    if (rc != VGX_OK) return rc;
    VGX_CUDA(c, cudaStreamSynchronize(c->stream));
    vgx_points_draw(p, (int)K, (int32_t*)c->h_pinned);
    float* d_pts = (float*)((char*)c->d_scratch + 9 * Kp * sizeof(double));
    int32_t* d_idx = (int32_t*)(d_pts + units * VGX_PT_UNIT_FLOATS);
    VGX_CUDA(c, cudaMemcpyAsync(d_idx, c->h_pinned, K * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    vgx_launch_reg_gather_samples(c->stream, p.data, d_idx, (int)K, d_pts);
    c->launches++;
    C.pts = d_pts;
  }
  RegPoseConst P;
  vgx_reg_pose_setup<TrigHostLibm>(ref_pose, read_pose, P);
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
