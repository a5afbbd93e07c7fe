// ESDF generation for a finished submap from its resident TSDF bricks:
//   cblox::TsdfEsdfSubmap::generateEsdf, called by VoxgraphSubmap::finishSubmap
//   (voxgraph/src/frontend/submap_collection/voxgraph_submap.cpp:86) =
//   voxblox::EsdfIntegrator::updateFromTsdfLayerBatch (upstream, absent here; restated from the
//   published algorithm as its fixed point; the CPU checker restates the same definition).
// The ESDF feeds the reference's default registration branch (use_esdf_distance = true,
// registration_cost_function.h:35; cpp:133-140) and findRelevantVoxelIndices' ESDF distance.
//
// B200 form: voxblox pops a bucketed priority queue on one thread; here every observed, non-fixed
// voxel repeatedly PULLS  d <- min(d, d_neighbour + |offset| voxel_size)  over its 26 neighbours
// (max on the negative side) in place until a sweep changes nothing.  The update is monotone, so
// any interleaving reaches the same least fixed point = the exact shortest quasi-Euclidean path in
// float arithmetic; max_distance / voxel_size sweeps (10-20) over the bricks suffice.
// Compiled with -fmad=false.
#include <math.h>
#include <string.h>

#include <vector>

#include "vgx_internal.h"

struct EsdfParams {
  float voxel_size, max_distance, default_distance, min_distance, min_weight;
  int vps, sh, n_blocks;
};

// (distance, observed) from the TSDF brick + "fixed" mask
__global__ void __launch_bounds__(256)
esdf_init_kernel(EsdfParams P, const float2* __restrict__ tsdf, float2* __restrict__ esdf,
                 unsigned char* __restrict__ fixed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nvox = (size_t)P.n_blocks << (3 * P.sh);
  if (i >= nvox) return;
  const float2 t = tsdf[i];
  float2 e = make_float2(0.f, 0.f);
  unsigned char f = 0;
  if (!(t.y < P.min_weight)) {
    e.y = 1.f;
    if (fabsf(t.x) < P.min_distance) { f = 1; e.x = t.x; }
    else e.x = (t.x > 0.f ? 1.f : -1.f) * P.default_distance;
  }
  esdf[i] = e;
  fixed[i] = f;
}

// 27 neighbour slots per block (centre included), -1 where the block does not exist
__global__ void esdf_neighbours_kernel(VgxHash hash, const int32_t* __restrict__ block_idx, int n_blocks,
                                       int* __restrict__ nb) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 27 * n_blocks) return;
  const int b = t / 27, k = t % 27;
  nb[t] = vgx_hash_find(hash, block_idx[3 * b] + (k % 3) - 1, block_idx[3 * b + 1] + ((k / 3) % 3) - 1,
                        block_idx[3 * b + 2] + (k / 9) - 1);
}

__global__ void __launch_bounds__(256)
esdf_sweep_kernel(EsdfParams P, float2* esdf, const unsigned char* __restrict__ fixed,
                  const int* __restrict__ nb, int* __restrict__ changed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nvox = (size_t)P.n_blocks << (3 * P.sh);
  if (i >= nvox) return;
  volatile float2* ev = esdf;
  const float2 me = make_float2(ev[i].x, ev[i].y);
  if (me.y == 0.f || fixed[i]) return;
  const int b = (int)(i >> (3 * P.sh));
  const int lin = (int)(i & (((size_t)1 << (3 * P.sh)) - 1));
  const int v0 = lin & (P.vps - 1), v1 = (lin >> P.sh) & (P.vps - 1), v2 = lin >> (2 * P.sh);
  const float s1 = 1.0f * P.voxel_size, s2 = sqrtf(2.0f) * P.voxel_size, s3 = sqrtf(3.0f) * P.voxel_size;
  float d = me.x;
#pragma unroll 1
  for (int dz = -1; dz <= 1; ++dz)
#pragma unroll 1
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        if (!dx && !dy && !dz) continue;
        int q0 = v0 + dx, q1 = v1 + dy, q2 = v2 + dz, bo0 = 1, bo1 = 1, bo2 = 1;
        if (q0 < 0) { q0 += P.vps; bo0 = 0; } else if (q0 >= P.vps) { q0 -= P.vps; bo0 = 2; }
        if (q1 < 0) { q1 += P.vps; bo1 = 0; } else if (q1 >= P.vps) { q1 -= P.vps; bo1 = 2; }
        if (q2 < 0) { q2 += P.vps; bo2 = 0; } else if (q2 >= P.vps) { q2 -= P.vps; bo2 = 2; }
        const int s = nb[27 * b + bo0 + 3 * bo1 + 9 * bo2];
        if (s < 0) continue;
        const size_t j = ((size_t)s << (3 * P.sh)) + q0 + (q1 << P.sh) + (q2 << (2 * P.sh));
        const float ny = ev[j].y;
        if (ny == 0.f) continue;
        const float dn = ev[j].x;
        const int m = (dx != 0) + (dy != 0) + (dz != 0);
        const float step = m == 1 ? s1 : (m == 2 ? s2 : s3);
        if (d > 0.f && dn > 0.f) {
          const float cand = dn + step;
          if (cand < P.max_distance && cand < d) d = cand;
        } else if (d < 0.f && dn < 0.f) {
          const float cand = dn - step;
          if (cand > -P.max_distance && cand > d) d = cand;
        }
      }
  if (d != me.x) {
    ev[i].x = d;
    *changed = 1;
  }
}

extern "C" void vgx_esdf_config_default(vgx_esdf_config* c) {
  if (!c) return;
  c->max_distance_m = 2.0f;       // voxblox::EsdfIntegrator::Config defaults
  c->default_distance_m = 2.0f;
  c->min_distance_m = 0.2f;
  c->min_weight = 1e-6f;
}

extern "C" int vgx_submap_generate_esdf(vgx_ctx* c, uint32_t id, const vgx_esdf_config* cfg, int* sweeps_out) {
  if (!c) return VGX_ERR_INVALID;
  VgxSubmap* s = c->find(id);
  if (!s) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_submap_generate_esdf: unknown submap");
  if (!s->finished) VGX_FAIL(c, VGX_ERR_INVALID, "vgx_submap_generate_esdf: submap is not finished");
  VGX_CUDA(c, cudaSetDevice(c->device));
  vgx_esdf_config dc;
  vgx_esdf_config_default(&dc);
  if (cfg) dc = *cfg;
  if (!(dc.max_distance_m > 0) || !(dc.min_distance_m >= 0))
    VGX_FAIL(c, VGX_ERR_INVALID, "vgx_submap_generate_esdf: invalid distances");
  vgx_graph_invalidate_registration(c);
  cudaFree(s->d_esdf); s->d_esdf = nullptr;
  cudaFree(s->d_view_esdf); s->d_view_esdf = nullptr;
  const size_t nvox = (size_t)s->n_blocks * s->vox_per_block;
  if (sweeps_out) *sweeps_out = 0;
  if (nvox == 0) {
    VGX_CUDA(c, cudaMalloc((void**)&s->d_view_esdf, 32));
    return VGX_OK;
  }
  EsdfParams P;
  P.voxel_size = s->voxel_size;
  P.max_distance = dc.max_distance_m; P.default_distance = dc.default_distance_m;
  P.min_distance = dc.min_distance_m; P.min_weight = dc.min_weight;
  P.vps = s->vps;
  P.sh = 0;
  while ((1 << P.sh) < s->vps) P.sh++;
  P.n_blocks = s->n_blocks;
  cudaStream_t st = c->stream;
  VGX_CUDA(c, cudaMalloc((void**)&s->d_esdf, sizeof(float2) * nvox));
  // scratch: fixed mask | neighbour table | changed flag
  const size_t fixed_b = (nvox + 255) & ~(size_t)255;
  const size_t nb_b = ((size_t)27 * s->n_blocks * sizeof(int) + 255) & ~(size_t)255;
  int rc = c->ensure_sort(fixed_b + nb_b + 256);
  if (rc != VGX_OK) return rc;
  unsigned char* d_fixed = (unsigned char*)c->d_sort;
  int* d_nb = (int*)((char*)c->d_sort + fixed_b);
  int* d_changed = (int*)((char*)c->d_sort + fixed_b + nb_b);
  const unsigned grid = (unsigned)((nvox + 255) / 256);
  esdf_init_kernel<<<grid, 256, 0, st>>>(P, s->d_dw, s->d_esdf, d_fixed);
  esdf_neighbours_kernel<<<(27 * s->n_blocks + 127) / 128, 128, 0, st>>>(s->hash, s->d_block_idx, s->n_blocks, d_nb);
  c->launches += 2;
  // relax to convergence: the flag is read back every other sweep
  int sweeps = 0;
  const int max_sweeps = 4 * (int)ceilf(dc.max_distance_m / s->voxel_size) + 16;
  for (;;) {
    VGX_CUDA(c, cudaMemsetAsync(d_changed, 0, sizeof(int), st));
    for (int k = 0; k < 2; ++k) {
      esdf_sweep_kernel<<<grid, 256, 0, st>>>(P, s->d_esdf, d_fixed, d_nb, d_changed);
      c->launches++;
      ++sweeps;
    }
    int changed = 0;
    VGX_CUDA(c, cudaMemcpyAsync(&changed, d_changed, sizeof(int), cudaMemcpyDeviceToHost, st));
    VGX_CUDA(c, cudaStreamSynchronize(st));
    if (!changed) break;
    if (sweeps > max_sweeps) VGX_FAIL(c, VGX_ERR_INVALID, "vgx_submap_generate_esdf: relaxation did not converge");
  }
  VGX_CUDA(c, cudaGetLastError());
  if (sweeps_out) *sweeps_out = sweeps;
  // registration view of the ESDF layer (observed <=> weight 1), same octet layout as the TSDF view
  VGX_CUDA(c, cudaMalloc((void**)&s->d_view_esdf, sizeof(float) * 8 * nvox));
  rc = vgx_submap_build_view(c, s, s->d_esdf, s->d_view_esdf);
  if (rc != VGX_OK) return rc;
  VGX_CUDA(c, cudaStreamSynchronize(st));
  return VGX_OK;
}

extern "C" int vgx_submap_download_esdf(vgx_ctx* c, uint32_t id, int max_blocks, float* distance,
                                        float* observed, int* n_out) {
  if (!c) return VGX_ERR_INVALID;
  VgxSubmap* s = c->find(id);
  if (!s) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_submap_download_esdf: unknown submap");
  if (!s->d_esdf && s->n_blocks > 0) VGX_FAIL(c, VGX_ERR_INVALID, "vgx_submap_download_esdf: no ESDF generated");
  VGX_CUDA(c, cudaSetDevice(c->device));
  const int n = s->n_blocks;
  if (n_out) *n_out = n;
  if (n > max_blocks) VGX_FAIL(c, VGX_ERR_CAPACITY, "vgx_submap_download_esdf: max_blocks too small");
  if (n == 0) return VGX_OK;
  const size_t nvox = (size_t)n * s->vox_per_block;
  std::vector<float2> h(nvox);
  VGX_CUDA(c, cudaMemcpyAsync(h.data(), s->d_esdf, sizeof(float2) * nvox, cudaMemcpyDeviceToHost, c->stream));
  VGX_CUDA(c, cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i < nvox; ++i) {
    if (distance) distance[i] = h[i].x;
    if (observed) observed[i] = h[i].y;
  }
  return VGX_OK;
}
