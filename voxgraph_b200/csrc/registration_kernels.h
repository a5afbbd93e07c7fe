// Host-visible declarations of the registration kernels (registration.cu).
#pragma once

#include "vgx_internal.h"

struct RegConstraintDev;
struct RegPoseConst;

#ifndef VGX_REG_THREADS
#define VGX_REG_THREADS 128        // 4 warps per CTA, each with its own TMA ring + software pipeline
#endif
#ifndef VGX_REG_MIN_BLOCKS
#define VGX_REG_MIN_BLOCKS 5       // resident CTAs per SM the register budget is sized for
#endif
#ifndef VGX_REG_TILE_UNITS
#define VGX_REG_TILE_UNITS 4       // units (x 32 points, x 640 B) per tile = per ticket = per TMA bulk copy
#endif
#define VGX_REG_UNIT 32            // points per unit = one warp iteration; tiles are cut on unit boundaries
#ifndef VGX_REG_STREAM_OCTETS
#define VGX_REG_STREAM_OCTETS 0   // ld.global.cs for the octet fetch
#endif
#define VGX_REG_NSUM 21            // 15 (upper 5x5) + 5 (gradient) + 1 (cost)
#define VGX_REG_NSTRIDE 24

struct RegTile {
  int constraint;
  int start;
  int count;
  int pad;
};

void vgx_launch_reg_pose_setup(cudaStream_t st, const RegConstraintDev* cons, const double* x,
                               RegPoseConst* poses, int n);
// Persistent warps draw tile tickets -> partial sums -> (last tile of each constraint)
// per-constraint sums csum[c][21].  sched: 2 ints, zero before the first launch (the kernel re-arms it).
void vgx_launch_reg_reduce(cudaStream_t st, const RegConstraintDev* cons, const RegPoseConst* poses,
                           const RegTile* tiles, int n_tiles, int n_ctas, const int* tile_begin,
                           int* counters, int* sched, double* partials, double* csum, bool jacobian);
// persistent grid size: SMs x co-resident CTAs
int vgx_reg_resident_ctas(int device);
// Fills the descriptor of one (reference -> reading) residual block.  Deterministic mode: pts / n /
// factor describe all registration points.  Sampling mode (*sampled = true): n = int(ratio * K),
// factor = 1 (all weights forced to 1) and pts is left null for the caller, who draws the
// indices (vgx_points_draw) and gathers them (vgx_launch_reg_gather_samples).
int vgx_fill_constraint(vgx_ctx* c, uint32_t ref_id, uint32_t read_id, const vgx_reg_config* cfg,
                        RegConstraintDev* out, bool* sampled = nullptr);
// dst (unit-major, ceil(n/32) units) <- the n points of src selected by d_idx, weight forced to 1
void vgx_launch_reg_gather_samples(cudaStream_t st, const float* src, const int32_t* d_idx, int n,
                                   float* dst);
