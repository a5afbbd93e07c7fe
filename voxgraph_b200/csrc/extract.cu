// Registration-point extraction from resident bricks:
//   VoxgraphSubmap::findRelevantVoxelIndices  voxgraph/src/frontend/submap_collection/voxgraph_submap.cpp:144-201
//   VoxgraphSubmap::findIsosurfaceVertices    voxgraph_submap.cpp:203-243
//   VoxgraphSubmap::getSubmapFrameSurfaceObb  voxgraph_submap.cpp:280-324
// so that a submap integrated on the GPU becomes a registration constraint's reference without its
// voxels ever leaving the device (finishSubmap, voxgraph_submap.cpp:84-107).
//
// findIsosurfaceVertices = voxblox MeshIntegrator::generateMesh + MeshLayer::getConnectedMesh +
// Interpolator::getVoxel.  Only the mesh VERTICES are used by voxgraph, and the set of vertices of
// a marching-cubes mesh is table independent: one vertex per cube edge whose end points differ in
// sign, for every cube whose 8 corners have weight > min_weight.  The connected mesh keeps the
// first vertex of each round(vertex / (0.5 voxel)) bucket; "first" is fixed canonically (blocks in
// allocation order, voxels in linear order, cube edges 0..11 - the reference walks a hash map).
// Data-parallel form: count -> scan -> emit candidates in canonical order -> bucket table with
// atomicMin(candidate index) -> the winners are interpolated -> scan -> compact.
// Compiled with -fmad=false (float expressions restated operation by operation, see oracle).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "vgx_internal.h"

// ------------------------------------------------------------------ exclusive scan (hand-written)
// out[i] = sum_{j<i} in[j] for i in [0, n]; out has n + 1 entries (out[n] = total).
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_kernel(const unsigned* __restrict__ in, unsigned* __restrict__ out, unsigned* __restrict__ tile_sums,
                 size_t n) {
  __shared__ unsigned s_warp[SCAN_THREADS / 32];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  unsigned v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    sum += v[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned w = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0u;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, w, off);
      if (lane >= off) w += t;
    }
    if (lane < SCAN_THREADS / 32) s_warp[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  unsigned run = incl - sum + (warp > 0 ? s_warp[warp - 1] : 0u);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == SCAN_THREADS - 1) tile_sums[blockIdx.x] = run;
}

// single CTA: exclusive scan of the tile sums in place, chunk by chunk with a running carry
__global__ void __launch_bounds__(1024) scan_sums_kernel(unsigned* __restrict__ sums, size_t m,
                                                          unsigned* __restrict__ total) {
  __shared__ unsigned s_warp[32];
  __shared__ unsigned s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (size_t c0 = 0; c0 < m; c0 += 1024) {
    const size_t i = c0 + threadIdx.x;
    const unsigned v = i < m ? sums[i] : 0u;
    unsigned incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned w = s_warp[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, w, off);
        if (lane >= off) w += t;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const unsigned carry = s_carry;
    const unsigned excl = carry + incl - v + (warp > 0 ? s_warp[warp - 1] : 0u);
    if (i < m) sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(unsigned* __restrict__ out, const unsigned* __restrict__ tile_sums, size_t n,
                const unsigned* __restrict__ total) {
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  const unsigned add = tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (base + k < n) out[base + k] += add;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total;
}

// d_tmp: (tiles + 1) unsigned of scratch.  3 launches.
int vgx_exclusive_scan_u32(vgx_ctx* c, const unsigned* d_in, unsigned* d_out, size_t n, unsigned* d_tmp) {
  if (n == 0) {
    VGX_CUDA(c, cudaMemsetAsync(d_out, 0, sizeof(unsigned), c->stream));
    return VGX_OK;
  }
  const size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  scan_tile_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, c->stream>>>(d_in, d_out, d_tmp, n);
  scan_sums_kernel<<<1, 1024, 0, c->stream>>>(d_tmp, tiles, d_tmp + tiles);
  scan_add_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, c->stream>>>(d_out, d_tmp, n, d_tmp + tiles);
  c->launches += 3;
  VGX_CUDA(c, cudaGetLastError());
  return VGX_OK;
}
size_t vgx_scan_tmp_count(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 2; }

// ------------------------------------------------------------------ layer access on the device
struct LayerDev {
  VgxHash hash;
  const float2* dw;
  const int32_t* block_idx;
  float voxel_size, voxel_size_inv, block_size, block_size_inv;
  int vps, sh, n_blocks;
};

__device__ __forceinline__ int ex_floor(float v) { return __float2int_rd(v); }

// Interpolator::getVoxel(pos, &voxel, interpolate = true): same steps as vgx_locate (registration.cuh)
// on the raw bricks, distance AND weight.  Restates oracle vgo_interp_voxel operation by operation.
__device__ bool ex_interp_voxel(const LayerDev& L, float p0, float p1, float p2, float& out_d, float& out_w) {
  const int vps = L.vps;
  const float p[3] = {p0, p1, p2};
  int b[3], v[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) b[a] = ex_floor(p[a] * L.block_size_inv + 1e-6f);
  if (vgx_hash_find(L.hash, b[0], b[1], b[2]) < 0) return false;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float origin = (float)b[a] * L.block_size;
    int vi = ex_floor((p[a] - origin) * L.voxel_size_inv + 1e-6f);
    vi = max(min(vi, vps - 1), 0);
    const float centre = origin + ((float)vi + 0.5f) * L.voxel_size;
    if (p[a] - centre < 0) {
      if (--vi < 0) { --b[a]; vi += vps; }
    }
    v[a] = vi;
  }
  float d[8], w[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int cv[3] = {v[0] + ((i >> 2) & 1), v[1] + ((i >> 1) & 1), v[2] + (i & 1)};
    int cb[3] = {b[0], b[1], b[2]};
    if (vgx_hash_find(L.hash, b[0], b[1], b[2]) < 0) return false;
#pragma unroll
    for (int a = 0; a < 3; ++a)
      if (cv[a] >= vps) { cb[a]++; cv[a] -= vps; }
    const int slot = vgx_hash_find(L.hash, cb[0], cb[1], cb[2]);
    if (slot < 0) return false;
    if (i == 0) {
      float off[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float centre = (float)b[a] * L.block_size + ((float)cv[a] + 0.5f) * L.voxel_size;
        off[a] = (p[a] - centre) * L.voxel_size_inv;
      }
      q[0] = 1.0f; q[1] = off[0]; q[2] = off[1]; q[3] = off[2];
      q[4] = off[0] * off[1]; q[5] = off[1] * off[2]; q[6] = off[2] * off[0];
      q[7] = off[0] * off[1] * off[2];
    }
    const float2 x = L.dw[((size_t)slot << (3 * L.sh)) + cv[0] + (cv[1] << L.sh) + (cv[2] << (2 * L.sh))];
    d[i] = x.x;
    w[i] = x.y;
    if (!(x.y > 1e-6f)) return false;
  }
  float res[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float* x = k == 0 ? d : w;
    float a[8];
    a[0] = x[0];
    a[1] = -x[0] + x[4];
    a[2] = -x[0] + x[2];
    a[3] = -x[0] + x[1];
    a[4] = x[0] - x[2] - x[4] + x[6];
    a[5] = x[0] - x[1] - x[2] + x[3];
    a[6] = x[0] - x[1] - x[4] + x[5];
    a[7] = -x[0] + x[1] + x[2] - x[3] + x[4] - x[5] - x[6] + x[7];
    float r = q[0] * a[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) r = r + q[i] * a[i];
    res[k] = r;
  }
  out_d = res[0];
  out_w = res[1];
  return true;
}

// ------------------------------------------------------------------ relevant voxels (cpp:144-201)
__global__ void __launch_bounds__(256)
relevant_count_kernel(LayerDev L, double min_w, double max_d, unsigned* __restrict__ counts) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nvox = (size_t)L.n_blocks << (3 * L.sh);
  if (i >= nvox) return;
  const float2 x = L.dw[i];
  counts[i] = ((double)x.y > min_w && fabs((double)x.x) < max_d) ? 1u : 0u;  // cpp:177-178
}

__device__ __forceinline__ float atomicMinFloat(float* addr, float v) {
  return (v >= 0) ? __int_as_float(atomicMin((int*)addr, __float_as_int(v)))
                  : __uint_as_float(atomicMax((unsigned*)addr, __float_as_uint(v)));
}
__device__ __forceinline__ float atomicMaxFloat(float* addr, float v) {
  return (v >= 0) ? __int_as_float(atomicMax((int*)addr, __float_as_int(v)))
                  : __uint_as_float(atomicMin((unsigned*)addr, __float_as_uint(v)));
}

// scatter into the unit-major point layout + surface OBB (cpp:280-324: centre -+ half voxel)
__global__ void __launch_bounds__(256)
relevant_emit_kernel(LayerDev L, const unsigned* __restrict__ counts, const unsigned* __restrict__ offsets,
                     const float2* __restrict__ esdf /* null: TSDF distance */, float* __restrict__ pts,
                     float* __restrict__ obb /* min[3], max[3] */) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nvox = (size_t)L.n_blocks << (3 * L.sh);
  if (i >= nvox || counts[i] == 0) return;
  const int slot = (int)(i >> (3 * L.sh));
  const int lin = (int)(i & (((size_t)1 << (3 * L.sh)) - 1));
  const int v[3] = {lin & (L.vps - 1), (lin >> L.sh) & (L.vps - 1), lin >> (2 * L.sh)};
  const float2 x = L.dw[i];
  float cpos[3];
  const float half = 0.5f * L.voxel_size;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    // Block::computeCoordinatesFromLinearIndex: origin + (voxel_index + 0.5) * voxel_size
    const float origin = (float)L.block_idx[3 * slot + a] * L.block_size;
    cpos[a] = origin + ((float)v[a] + 0.5f) * L.voxel_size;
    atomicMinFloat(obb + a, cpos[a] - half);
    atomicMaxFloat(obb + 3 + a, cpos[a] + half);
  }
  float* u = pts + vgx_pt_index((size_t)offsets[i], 0);
  // cpp:183-191: the ESDF distance when use_esdf_distance, the weight is always the TSDF weight
  u[0] = cpos[0]; u[32] = cpos[1]; u[64] = cpos[2]; u[96] = esdf ? esdf[i].x : x.x; u[128] = x.y;
}

// ------------------------------------------------------------------ isosurface vertices (cpp:203-243)
__constant__ int c_cube_off[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                                     {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
__constant__ int c_edge[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6},
                                  {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};

// gathers the cube of voxel i: returns false unless all 8 corners have weight > min_weight
__device__ __forceinline__ bool ex_cube(const LayerDev& L, size_t i, float min_weight, float sdf[8],
                                        float coords[3]) {
  const int slot = (int)(i >> (3 * L.sh));
  const int lin = (int)(i & (((size_t)1 << (3 * L.sh)) - 1));
  const int v[3] = {lin & (L.vps - 1), (lin >> L.sh) & (L.vps - 1), lin >> (2 * L.sh)};
  const int b[3] = {L.block_idx[3 * slot], L.block_idx[3 * slot + 1], L.block_idx[3 * slot + 2]};
#pragma unroll
  for (int a = 0; a < 3; ++a) coords[a] = (float)b[a] * L.block_size + ((float)v[a] + 0.5f) * L.voxel_size;
  int nslot[8];
  nslot[0] = slot;
#pragma unroll
  for (int k = 1; k < 8; ++k) {  // k bit0 = x overflow, bit1 = y, bit2 = z
    const bool ux = (k & 1) && v[0] == L.vps - 1, uy = (k & 2) && v[1] == L.vps - 1,
               uz = (k & 4) && v[2] == L.vps - 1;
    const int j = (ux ? 1 : 0) | (uy ? 2 : 0) | (uz ? 4 : 0);
    nslot[k] = (j == k) ? vgx_hash_find(L.hash, b[0] + (ux ? 1 : 0), b[1] + (uy ? 1 : 0), b[2] + (uz ? 1 : 0))
                        : nslot[j];
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int ox = c_cube_off[c][0], oy = c_cube_off[c][1], oz = c_cube_off[c][2];
    const int s = nslot[ox | (oy << 1) | (oz << 2)];
    if (s < 0) return false;
    const int x = (v[0] + ox) & (L.vps - 1), y = (v[1] + oy) & (L.vps - 1), z = (v[2] + oz) & (L.vps - 1);
    const float2 q = L.dw[((size_t)s << (3 * L.sh)) + x + (y << L.sh) + (z << (2 * L.sh))];
    if (q.y <= min_weight) return false;  // utils::getSdfIfValid
    sdf[c] = q.x;
  }
  return true;
}

__device__ __forceinline__ bool ex_edge_crosses(float s1, float s2) {
  return (s1 < 0 && s2 >= 0) || (s1 >= 0 && s2 < 0);
}

__global__ void __launch_bounds__(128)
iso_count_kernel(LayerDev L, float min_weight, unsigned* __restrict__ counts) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nvox = (size_t)L.n_blocks << (3 * L.sh);
  if (i >= nvox) return;
  float sdf[8], coords[3];
  unsigned n = 0;
  if (ex_cube(L, i, min_weight, sdf, coords)) {
#pragma unroll
    for (int e = 0; e < 12; ++e) n += ex_edge_crosses(sdf[c_edge[e][0]], sdf[c_edge[e][1]]) ? 1u : 0u;
  }
  counts[i] = n;
}

struct __align__(16) IsoBucket {
  unsigned long long key;   // packed bucket coordinates, VGX_EMPTY_KEY when free
  unsigned first;           // smallest candidate index seen
  unsigned pad;
};

__global__ void iso_table_clear_kernel(IsoBucket* t, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { t[i].key = VGX_EMPTY_KEY; t[i].first = 0xFFFFFFFFu; t[i].pad = 0; }
}

__device__ __forceinline__ unsigned long long iso_bucket_key(const float p[3], double threshold_inv) {
  // createConnectedMesh: round(double(vertex) * threshold_inv) per axis, 21 bits each (biased)
  unsigned long long k = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const long long r = llround((double)p[a] * threshold_inv);
    k |= ((unsigned long long)((r + (1ll << 20)) & 0x1FFFFF)) << (21 * a);
  }
  return k;
}

__device__ __forceinline__ size_t iso_slot_of(unsigned long long key, size_t mask) {
  unsigned long long h = key * 0x9E3779B97F4A7C15ull;
  h ^= h >> 29;
  return (size_t)h & mask;
}

// candidates in canonical order; each registers itself in its bucket (atomicMin of its index)
__global__ void __launch_bounds__(128)
iso_emit_kernel(LayerDev L, float min_weight, double threshold_inv, const unsigned* __restrict__ counts,
                const unsigned* __restrict__ offsets, float* __restrict__ cand /* M x 3 */,
                unsigned long long* __restrict__ cand_key, IsoBucket* __restrict__ table, size_t mask) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nvox = (size_t)L.n_blocks << (3 * L.sh);
  if (i >= nvox || counts[i] == 0) return;
  float sdf[8], coords[3];
  if (!ex_cube(L, i, min_weight, sdf, coords)) return;
  unsigned o = offsets[i];
#pragma unroll
  for (int e = 0; e < 12; ++e) {
    const int e0 = c_edge[e][0], e1 = c_edge[e][1];
    const float s1 = sdf[e0], s2 = sdf[e1];
    if (!ex_edge_crosses(s1, s2)) continue;
    float p[3];
    const float diff = s1 - s2;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      // corner_coords = coords + cube_coord_offsets (index offset * voxel_size)
      const float c0 = coords[a] + (float)c_cube_off[e0][a] * L.voxel_size;
      const float c1 = coords[a] + (float)c_cube_off[e1][a] * L.voxel_size;
      if (fabsf(diff) >= 1e-6f) {
        const float t = s1 / diff;
        p[a] = c0 + t * (c1 - c0);
      } else {
        p[a] = 0.5f * (c0 + c1);
      }
    }
    cand[3 * (size_t)o] = p[0]; cand[3 * (size_t)o + 1] = p[1]; cand[3 * (size_t)o + 2] = p[2];
    const unsigned long long key = iso_bucket_key(p, threshold_inv);
    cand_key[o] = key;
    size_t h = iso_slot_of(key, mask);
    for (;;) {
      const unsigned long long prev = atomicCAS(&table[h].key, (unsigned long long)VGX_EMPTY_KEY, key);
      if (prev == VGX_EMPTY_KEY || prev == key) { atomicMin(&table[h].first, o); break; }
      h = (h + 1) & mask;
    }
    ++o;
  }
}

// the first candidate of every bucket is a connected-mesh vertex; interpolate (distance, weight)
__global__ void __launch_bounds__(128)
iso_select_kernel(LayerDev L, const float* __restrict__ cand, const unsigned long long* __restrict__ cand_key,
                  const IsoBucket* __restrict__ table, size_t mask, unsigned M, unsigned* __restrict__ keep,
                  float2* __restrict__ cand_dw) {
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= M) return;
  const unsigned long long key = cand_key[c];
  size_t h = iso_slot_of(key, mask);
  while (table[h].key != key) h = (h + 1) & mask;
  unsigned k = 0;
  if (table[h].first == c) {
    float d, w;
    if (ex_interp_voxel(L, cand[3 * (size_t)c], cand[3 * (size_t)c + 1], cand[3 * (size_t)c + 2], d, w)) {
      k = 1;
      cand_dw[c] = make_float2(d, w);
    }
  }
  keep[c] = k;
}

__global__ void __launch_bounds__(128)
iso_compact_kernel(LayerDev L, const float* __restrict__ cand, const float2* __restrict__ cand_dw,
                   const unsigned* __restrict__ keep, const unsigned* __restrict__ offsets, unsigned M,
                   float* __restrict__ pts, unsigned* __restrict__ iso_block_flags) {
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= M || keep[c] == 0) return;
  const float p[3] = {cand[3 * (size_t)c], cand[3 * (size_t)c + 1], cand[3 * (size_t)c + 2]};
  float* u = pts + vgx_pt_index((size_t)offsets[c], 0);
  const float2 dw = cand_dw[c];
  u[0] = p[0]; u[32] = p[1]; u[64] = p[2]; u[96] = dw.x; u[128] = dw.y;
  // isosurface_blocks_ (cpp:237-240): computeBlockIndexFromCoordinates(vertex)
  const int b0 = ex_floor(p[0] * L.block_size_inv + 1e-6f), b1 = ex_floor(p[1] * L.block_size_inv + 1e-6f),
            b2 = ex_floor(p[2] * L.block_size_inv + 1e-6f);
  const int slot = vgx_hash_find(L.hash, b0, b1, b2);
  if (slot >= 0) iso_block_flags[slot] = 1u;
}

__global__ void gather_weights_kernel(const float* __restrict__ pts, int n, float* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w[i] = pts[vgx_pt_index((size_t)i, 4)];
}

__global__ void unpack_points_kernel(const float* __restrict__ pts, int n, float* __restrict__ xyz,
                                     float* __restrict__ d, float* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* u = pts + vgx_pt_index((size_t)i, 0);
  if (xyz) { xyz[3 * (size_t)i] = u[0]; xyz[3 * (size_t)i + 1] = u[32]; xyz[3 * (size_t)i + 2] = u[64]; }
  if (d) d[i] = u[96];
  if (w) w[i] = u[128];
}

// ------------------------------------------------------------------ host orchestration
static LayerDev layer_of(const VgxSubmap* s) {
  LayerDev L;
  L.hash = s->hash;
  L.dw = s->d_dw;
  L.block_idx = s->d_block_idx;
  L.voxel_size = s->voxel_size; L.voxel_size_inv = s->voxel_size_inv;
  L.block_size = s->block_size; L.block_size_inv = s->block_size_inv;
  L.vps = s->vps;
  L.sh = 0;
  while ((1 << L.sh) < s->vps) L.sh++;
  L.n_blocks = s->n_blocks;
  return L;
}

// takes ownership of d_pts (n points, unit-major, zero padded); weights come back to the host once
// for the sampler's cumulative sums and summed_reference_weight (cpp:124), in point order
static int install_points(vgx_ctx* c, VgxSubmap* s, int type, float* d_pts, int n) {
  VgxPoints& p = s->points[type];
  if (p.data) cudaFree(p.data);
  p = VgxPoints();
  if (n == 0) {
    cudaFree(d_pts);
    return VGX_OK;
  }
  std::vector<float> w((size_t)n);
  int rc = c->ensure_scratch(sizeof(float) * (size_t)n);
  if (rc != VGX_OK) { cudaFree(d_pts); return rc; }
  gather_weights_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_pts, n, (float*)c->d_scratch);
  c->launches++;
  cudaError_t e = cudaMemcpyAsync(w.data(), c->d_scratch, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    cudaFree(d_pts);
    c->set_error(std::string("registration point extraction: ") + cudaGetErrorString(e));
    return VGX_ERR_CUDA;
  }
  p.cumulative_w.resize((size_t)n);
  double sum = 0;
  for (int i = 0; i < n; ++i) { sum += (double)w[i]; p.cumulative_w[i] = sum; }
  p.n = n;
  p.data = d_pts;
  p.sum_w = sum;
  return VGX_OK;
}

static int alloc_points(vgx_ctx* c, int n, float** out) {
  *out = nullptr;
  const size_t units = ((size_t)std::max(n, 1) + VGX_PT_UNIT - 1) / VGX_PT_UNIT;
  VGX_CUDA(c, cudaMalloc((void**)out, units * VGX_PT_UNIT_FLOATS * sizeof(float)));
  VGX_CUDA(c, cudaMemsetAsync(*out, 0, units * VGX_PT_UNIT_FLOATS * sizeof(float), c->stream));
  return VGX_OK;
}

extern "C" void vgx_registration_filter_default(vgx_registration_filter* f) {
  if (!f) return;
  f->min_voxel_weight = 1.0;     // voxgraph_submap.h:27
  f->max_voxel_distance = 0.3;   // voxgraph_submap.h:28
  f->use_esdf_distance = 0;      // reference default true; the ESDF branch needs an ESDF layer
}

extern "C" int vgx_submap_extract_points(vgx_ctx* c, uint32_t id, const vgx_registration_filter* filter) {
  if (!c) return VGX_ERR_INVALID;
  VgxSubmap* s = c->find(id);
  if (!s) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_submap_extract_points: unknown submap");
  if (!s->finished) VGX_FAIL(c, VGX_ERR_INVALID, "vgx_submap_extract_points: submap is not finished");
  vgx_registration_filter f;
  vgx_registration_filter_default(&f);
  if (filter) f = *filter;
  if (f.use_esdf_distance && !s->d_esdf && s->n_blocks > 0)
    VGX_FAIL(c, VGX_ERR_INVALID, "vgx_submap_extract_points: use_esdf_distance needs the submap's ESDF "
                                 "(vgx_submap_generate_esdf first, as finishSubmap does)");
  VGX_CUDA(c, cudaSetDevice(c->device));
  vgx_graph_invalidate_registration(c);
  const LayerDev L = layer_of(s);
  const size_t nvox = (size_t)s->n_blocks * s->vox_per_block;
  s->surface_obb_valid = false;
  s->iso_blocks.clear();
  cudaFree(s->d_iso_idx);
  s->d_iso_idx = nullptr;
  s->n_iso = 0;
  s->points_extracted = true;
  if (nvox == 0) {
    int rc = install_points(c, s, VGX_POINTS_VOXELS, nullptr, 0);
    if (rc == VGX_OK) rc = install_points(c, s, VGX_POINTS_ISOSURFACE, nullptr, 0);
    return rc;
  }
  if (nvox >= 0xFFFFFFFFull / 12) VGX_FAIL(c, VGX_ERR_CAPACITY, "submap too large for 32-bit candidate indices");
  cudaStream_t st = c->stream;
  // scratch: counts | offsets (+1) | scan tmp | obb[6] | block flags
  const size_t cnt_b = ((nvox + 1) * sizeof(unsigned) + 255) & ~(size_t)255;
  const size_t tmp_b = (vgx_scan_tmp_count(nvox * 12) * sizeof(unsigned) + 255) & ~(size_t)255;
  const size_t flag_b = ((size_t)s->n_blocks * sizeof(unsigned) + 255) & ~(size_t)255;
  int rc = c->ensure_sort(2 * cnt_b + tmp_b + 256 + flag_b);
  if (rc != VGX_OK) return rc;
  char* base = (char*)c->d_sort;
  unsigned* d_counts = (unsigned*)base;
  unsigned* d_offsets = (unsigned*)(base + cnt_b);
  unsigned* d_tmp = (unsigned*)(base + 2 * cnt_b);
  float* d_obb = (float*)(base + 2 * cnt_b + tmp_b);
  unsigned* d_flags = (unsigned*)(base + 2 * cnt_b + tmp_b + 256);
  const unsigned grid256 = (unsigned)((nvox + 255) / 256), grid128 = (unsigned)((nvox + 127) / 128);

  // ---- relevant voxels + surface OBB
  {
    const float inf = INFINITY;
    const float init[6] = {inf, inf, inf, -inf, -inf, -inf};
    VGX_CUDA(c, cudaMemcpyAsync(d_obb, init, sizeof(init), cudaMemcpyHostToDevice, st));
    relevant_count_kernel<<<grid256, 256, 0, st>>>(L, f.min_voxel_weight, f.max_voxel_distance, d_counts);
    c->launches++;
    rc = vgx_exclusive_scan_u32(c, d_counts, d_offsets, nvox, d_tmp);
    if (rc != VGX_OK) return rc;
    unsigned total = 0;
    VGX_CUDA(c, cudaMemcpyAsync(&total, d_offsets + nvox, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    VGX_CUDA(c, cudaStreamSynchronize(st));
    float* d_pts = nullptr;
    rc = alloc_points(c, (int)total, &d_pts);
    if (rc != VGX_OK) return rc;
    relevant_emit_kernel<<<grid256, 256, 0, st>>>(L, d_counts, d_offsets, f.use_esdf_distance ? s->d_esdf : nullptr,
                                                 d_pts, d_obb);
    c->launches++;
    VGX_CUDA(c, cudaGetLastError());
    float obb[6];
    VGX_CUDA(c, cudaMemcpyAsync(obb, d_obb, sizeof(obb), cudaMemcpyDeviceToHost, st));
    VGX_CUDA(c, cudaStreamSynchronize(st));
    for (int a = 0; a < 3; ++a) { s->surface_obb_min[a] = obb[a]; s->surface_obb_max[a] = obb[3 + a]; }
    s->surface_obb_valid = total > 0;
    rc = install_points(c, s, VGX_POINTS_VOXELS, d_pts, (int)total);
    if (rc != VGX_OK) return rc;
  }
  // ---- isosurface vertices
  {
    const float min_weight = (float)f.min_voxel_weight;                 // cpp:211-212
    const float threshold = (float)(0.5 * (double)s->voxel_size);       // cpp:220
    const double threshold_inv = 1.0 / (double)threshold;               // createConnectedMesh
    iso_count_kernel<<<grid128, 128, 0, st>>>(L, min_weight, d_counts);
    c->launches++;
    rc = vgx_exclusive_scan_u32(c, d_counts, d_offsets, nvox, d_tmp);
    if (rc != VGX_OK) return rc;
    unsigned M = 0;
    VGX_CUDA(c, cudaMemcpyAsync(&M, d_offsets + nvox, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    VGX_CUDA(c, cudaStreamSynchronize(st));
    unsigned kept = 0;
    float* d_pts = nullptr;
    if (M > 0) {
      size_t tsize = 1024;
      while (tsize < 2 * (size_t)M) tsize <<= 1;
      const size_t cand_b = ((size_t)M * 3 * sizeof(float) + 255) & ~(size_t)255;
      const size_t key_b = ((size_t)M * sizeof(unsigned long long) + 255) & ~(size_t)255;
      const size_t keep_b = (((size_t)M + 1) * sizeof(unsigned) + 255) & ~(size_t)255;
      const size_t dw_b = ((size_t)M * sizeof(float2) + 255) & ~(size_t)255;
      const size_t tmp2_b = (vgx_scan_tmp_count(M) * sizeof(unsigned) + 255) & ~(size_t)255;
      rc = c->ensure_sort2(cand_b + key_b + 2 * keep_b + dw_b + tmp2_b + tsize * sizeof(IsoBucket));
      if (rc != VGX_OK) return rc;
      char* b2 = (char*)c->d_sort2;
      IsoBucket* d_table = (IsoBucket*)b2;
      float* d_cand = (float*)(b2 + tsize * sizeof(IsoBucket));
      unsigned long long* d_key = (unsigned long long*)((char*)d_cand + cand_b);
      unsigned* d_keep = (unsigned*)((char*)d_key + key_b);
      unsigned* d_koff = (unsigned*)((char*)d_keep + keep_b);
      float2* d_cdw = (float2*)((char*)d_koff + keep_b);
      unsigned* d_tmp2 = (unsigned*)((char*)d_cdw + dw_b);
      iso_table_clear_kernel<<<(unsigned)((tsize + 255) / 256), 256, 0, st>>>(d_table, tsize);
      iso_emit_kernel<<<grid128, 128, 0, st>>>(L, min_weight, threshold_inv, d_counts, d_offsets, d_cand, d_key,
                                              d_table, tsize - 1);
      iso_select_kernel<<<(M + 127) / 128, 128, 0, st>>>(L, d_cand, d_key, d_table, tsize - 1, M, d_keep, d_cdw);
      c->launches += 3;
      rc = vgx_exclusive_scan_u32(c, d_keep, d_koff, M, d_tmp2);
      if (rc != VGX_OK) return rc;
      VGX_CUDA(c, cudaMemcpyAsync(&kept, d_koff + M, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
      VGX_CUDA(c, cudaStreamSynchronize(st));
      rc = alloc_points(c, (int)kept, &d_pts);
      if (rc != VGX_OK) return rc;
      VGX_CUDA(c, cudaMemsetAsync(d_flags, 0, flag_b, st));
      iso_compact_kernel<<<(M + 127) / 128, 128, 0, st>>>(L, d_cand, d_cdw, d_keep, d_koff, M, d_pts, d_flags);
      c->launches++;
      VGX_CUDA(c, cudaGetLastError());
      // isosurface block list (slots, ascending) for overlapsWith
      std::vector<unsigned> flags((size_t)s->n_blocks);
      VGX_CUDA(c, cudaMemcpyAsync(flags.data(), d_flags, sizeof(unsigned) * (size_t)s->n_blocks,
                                  cudaMemcpyDeviceToHost, st));
      VGX_CUDA(c, cudaStreamSynchronize(st));
      for (int b = 0; b < s->n_blocks; ++b)
        if (flags[b]) s->iso_blocks.push_back(b);
      if (!s->iso_blocks.empty()) {
        std::vector<int32_t> bidx(3 * (size_t)s->n_blocks), iso(3 * s->iso_blocks.size());
        VGX_CUDA(c, cudaMemcpyAsync(bidx.data(), s->d_block_idx, sizeof(int32_t) * bidx.size(),
                                    cudaMemcpyDeviceToHost, st));
        VGX_CUDA(c, cudaStreamSynchronize(st));
        for (size_t k = 0; k < s->iso_blocks.size(); ++k)
          for (int a = 0; a < 3; ++a) iso[3 * k + a] = bidx[3 * (size_t)s->iso_blocks[k] + a];
        VGX_CUDA(c, cudaMalloc((void**)&s->d_iso_idx, sizeof(int32_t) * iso.size()));
        VGX_CUDA(c, cudaMemcpyAsync(s->d_iso_idx, iso.data(), sizeof(int32_t) * iso.size(),
                                    cudaMemcpyHostToDevice, st));
        VGX_CUDA(c, cudaStreamSynchronize(st));
        s->n_iso = (int)s->iso_blocks.size();
      }
    } else {
      rc = alloc_points(c, 0, &d_pts);
      if (rc != VGX_OK) return rc;
    }
    rc = install_points(c, s, VGX_POINTS_ISOSURFACE, d_pts, (int)kept);
    if (rc != VGX_OK) return rc;
  }
  return VGX_OK;
}

extern "C" int vgx_submap_finish_ex(vgx_ctx* c, uint32_t id, const vgx_registration_filter* filter) {
  int rc = vgx_submap_finish(c, id);
  if (rc != VGX_OK) return rc;
  return vgx_submap_extract_points(c, id, filter);
}

extern "C" int vgx_submap_num_points(vgx_ctx* c, uint32_t id, int type, int* n) {
  if (!c || !n || type < 0 || type > 1) return VGX_ERR_INVALID;
  VgxSubmap* s = c->find(id);
  if (!s) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_submap_num_points: unknown submap");
  *n = s->points[type].n;
  return VGX_OK;
}

extern "C" int vgx_submap_download_points(vgx_ctx* c, uint32_t id, int type, int max_n, float* xyz,
                                          float* distance, float* weight, int* n_out) {
  if (!c || type < 0 || type > 1) return VGX_ERR_INVALID;
  VgxSubmap* s = c->find(id);
  if (!s) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_submap_download_points: unknown submap");
  const VgxPoints& p = s->points[type];
  if (n_out) *n_out = p.n;
  if (p.n > max_n) VGX_FAIL(c, VGX_ERR_CAPACITY, "vgx_submap_download_points: max_n too small");
  if (p.n == 0) return VGX_OK;
  VGX_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)p.n;
  int rc = c->ensure_scratch(5 * n * sizeof(float));
  if (rc != VGX_OK) return rc;
  float* d_xyz = (float*)c->d_scratch;
  float* d_d = d_xyz + 3 * n;
  float* d_w = d_d + n;
  unpack_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(p.data, p.n, d_xyz, d_d, d_w);
  c->launches++;
  if (xyz) VGX_CUDA(c, cudaMemcpyAsync(xyz, d_xyz, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (distance) VGX_CUDA(c, cudaMemcpyAsync(distance, d_d, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (weight) VGX_CUDA(c, cudaMemcpyAsync(weight, d_w, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  VGX_CUDA(c, cudaStreamSynchronize(c->stream));
  return VGX_OK;
}

extern "C" int vgx_submap_surface_obb(vgx_ctx* c, uint32_t id, float obb_min[3], float obb_max[3]) {
  if (!c || !obb_min || !obb_max) return VGX_ERR_INVALID;
  VgxSubmap* s = c->find(id);
  if (!s) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_submap_surface_obb: unknown submap");
  if (!s->surface_obb_valid) return VGX_ZERO_WEIGHT;  // no voxel qualifies: the box stays +-inf
  for (int a = 0; a < 3; ++a) { obb_min[a] = s->surface_obb_min[a]; obb_max[a] = s->surface_obb_max[a]; }
  return VGX_OK;
}
