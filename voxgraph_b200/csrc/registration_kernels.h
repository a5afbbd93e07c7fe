// Host-visible declarations of the registration kernels (registration.cu).
#pragma once

#include "vgx_internal.h"

struct RegConstraintDev;
struct RegPoseConst;

#define VGX_REG_THREADS 256
#ifndef VGX_REG_MIN_BLOCKS
#define VGX_REG_MIN_BLOCKS 3       // resident CTAs per SM the register budget is sized for
#endif
#ifndef VGX_REG_PREFETCH
#define VGX_REG_PREFETCH 0         // software-pipelined point loads
#endif
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
// Persistent CTAs walk their tiles -> partial sums -> (last tile of each constraint)
// per-constraint sums csum[c][21].
void vgx_launch_reg_reduce(cudaStream_t st, const RegConstraintDev* cons, const RegPoseConst* poses,
                           const RegTile* tiles, int n_ctas, const int* cta_tile_begin,
                           const int* tile_begin, int* counters, double* partials, double* csum,
                           int grid_capacity, bool jacobian);
int vgx_reg_resident_ctas(int device);
int vgx_fill_constraint(vgx_ctx* c, uint32_t ref_id, uint32_t read_id, const vgx_reg_config* cfg,
                        RegConstraintDev* out);
