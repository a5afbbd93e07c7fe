// Context + brick store: voxblox::Layer/Block/AnyIndexHash replacement in HBM.
// Reference semantics: SURVEY.md Appendix A.2 (voxblox core) as used by
// voxgraph/src/frontend/submap_collection/voxgraph_submap.cpp.
#include <math.h>
#include <string.h>

#include <limits.h>

#include <algorithm>
#include <new>
#include <vector>

#include "vgx_internal.h"

// ------------------------------------------------------------------ context
int vgx_ctx::ensure_scratch(size_t bytes) {
  if (bytes <= scratch_bytes) return VGX_OK;
  if (d_scratch) cudaFree(d_scratch);
  d_scratch = nullptr;
  scratch_bytes = 0;
  size_t want = bytes + bytes / 4;
  VGX_CUDA(this, cudaMalloc(&d_scratch, want));
  scratch_bytes = want;
  return VGX_OK;
}

int vgx_ctx::ensure_pinned(size_t bytes) {
  if (bytes <= pinned_bytes) return VGX_OK;
  if (h_pinned) cudaFreeHost(h_pinned);
  h_pinned = nullptr;
  pinned_bytes = 0;
  size_t want = bytes + bytes / 4;
  VGX_CUDA(this, cudaMallocHost(&h_pinned, want));
  pinned_bytes = want;
  return VGX_OK;
}

static int grow(vgx_ctx* c, void** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap) return VGX_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  const size_t want = bytes + bytes / 4;
  VGX_CUDA(c, cudaMalloc(p, want));
  *cap = want;
  return VGX_OK;
}
int vgx_ctx::ensure_sort(size_t bytes) { return grow(this, &d_sort, &sort_bytes, bytes); }
int vgx_ctx::ensure_sort2(size_t bytes) { return grow(this, &d_sort2, &sort2_bytes, bytes); }

extern "C" int vgx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" int vgx_ctx_create(int device, vgx_ctx** out) {
  if (!out) return VGX_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n)
    return VGX_ERR_CUDA;  // no CPU fallback: the product path needs a GPU
  if (cudaSetDevice(device) != cudaSuccess) return VGX_ERR_CUDA;
  vgx_ctx* c = new (std::nothrow) vgx_ctx();
  if (!c) return VGX_ERR_NOMEM;
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
    delete c;
    return VGX_ERR_CUDA;
  }
  *out = c;
  return VGX_OK;
}

static void free_points(VgxPoints& p) {
  if (p.data) cudaFree(p.data);
  p = VgxPoints();
}

static void free_submap(VgxSubmap* s) {
  if (!s) return;
  cudaFree(s->hash.entries);
  cudaFree(s->d_block_idx);
  cudaFree(s->d_dw);
  cudaFree(s->d_view);
  cudaFree(s->d_esdf);
  cudaFree(s->d_view_esdf);
  cudaFree(s->d_counters);
  cudaFree(s->d_grid);
  cudaFree(s->d_grid16);
  cudaFree(s->d_iso_idx);
  free_points(s->points[0]);
  free_points(s->points[1]);
  delete s;
}

extern "C" void vgx_ctx_destroy(vgx_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  vgx_comm_destroy(c);
  vgx_p2p_free(c);
  vgx_graph_free(c);
  for (auto& kv : c->submaps) free_submap(kv.second);
  if (c->d_scratch) cudaFree(c->d_scratch);
  if (c->d_sort) cudaFree(c->d_sort);
  if (c->d_sort2) cudaFree(c->d_sort2);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" const char* vgx_last_error(const vgx_ctx* c) { return c ? c->error.c_str() : "null context"; }
extern "C" void* vgx_ctx_stream(vgx_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int vgx_ctx_synchronize(vgx_ctx* c) {
  if (!c) return VGX_ERR_INVALID;
  VGX_CUDA(c, cudaStreamSynchronize(c->stream));
  return VGX_OK;
}

extern "C" int vgx_profile_enable(vgx_ctx* c, int on) {
  if (!c) return VGX_ERR_INVALID;
  c->profile = on != 0;
  return VGX_OK;
}
extern "C" int vgx_profile_reset(vgx_ctx* c) {
  if (!c) return VGX_ERR_INVALID;
  for (auto& p : c->prof) p = VgxProfileSlot();
  c->launches = 0;
  return VGX_OK;
}
extern "C" int vgx_profile_get(vgx_ctx* c, int which, double* total_ms, int64_t* launches) {
  if (!c || which < 0 || which >= 10) return VGX_ERR_INVALID;
  if (total_ms) *total_ms = c->prof[which].total_ms;
  if (launches) *launches = c->prof[which].launches;
  return VGX_OK;
}
extern "C" int64_t vgx_launch_count(vgx_ctx* c) { return c ? c->launches : 0; }

// ------------------------------------------------------------------ kernels
__global__ void hash_clear_kernel(VgxHashEntry* entries, uint32_t size) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < size) {
    VgxHashEntry e;
    e.key = VGX_EMPTY_KEY;
    e.val = -1;
    e.pad = 0;
    entries[i] = e;
  }
}

// Insert n known-distinct blocks (slot = position in the list).
__global__ void hash_insert_kernel(VgxHash h, const int32_t* block_idx, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int bx = block_idx[3 * s], by = block_idx[3 * s + 1], bz = block_idx[3 * s + 2];
  const uint64_t key = vgx_pack_key(bx, by, bz);
  uint32_t i = vgx_hash_index(bx, by, bz, h.mask);
  for (;;) {
    unsigned long long prev =
        atomicCAS((unsigned long long*)&h.entries[i].key, (unsigned long long)VGX_EMPTY_KEY,
                  (unsigned long long)key);
    if (prev == VGX_EMPTY_KEY || prev == key) {
      h.entries[i].val = s;  // duplicates in the input: last writer wins
      return;
    }
    i = (i + 1) & h.mask;
  }
}

// (distance[], weight[]) -> interleaved float2 bricks.
__global__ void interleave_kernel(const float* __restrict__ d, const float* __restrict__ w,
                                  float2* __restrict__ dw, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dw[i] = make_float2(d[i], w[i]);
}

// Registration view: one thread per voxel gathers its 2x2x2 forward neighbourhood (across
// brick faces through the hash) and stores the 8 distances, NaN where
// utils::isObservedVoxel (weight > 1e-6) fails or the neighbour block does not exist.
__global__ void __launch_bounds__(256)
build_view_kernel(const float2* __restrict__ dw, const int32_t* __restrict__ block_idx, VgxHash hash,
                  float* __restrict__ view, int n_blocks, int vps, int sh) {
  const size_t vpb = (size_t)1 << (3 * sh);
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n_blocks * vpb) return;
  const int slot = (int)(i >> (3 * sh));
  const int lin = (int)(i & (vpb - 1));
  const int vx = lin & (vps - 1), vy = (lin >> sh) & (vps - 1), vz = lin >> (2 * sh);
  const int bx = block_idx[3 * slot], by = block_idx[3 * slot + 1], bz = block_idx[3 * slot + 2];
  const bool cx = vx == vps - 1, cy = vy == vps - 1, cz = vz == vps - 1;
  int s[8];
  s[0] = slot;
#pragma unroll
  for (int c = 1; c < 8; ++c) {
    const bool ux = (c & 4) && cx, uy = (c & 2) && cy, uz = (c & 1) && cz;
    const int j = (ux ? 4 : 0) | (uy ? 2 : 0) | (uz ? 1 : 0);
    if (j == c) s[c] = vgx_hash_find(hash, bx + (ux ? 1 : 0), by + (uy ? 1 : 0), bz + (uz ? 1 : 0));
    else s[c] = s[j];
  }
  float o[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int x = (vx + ((c >> 2) & 1)) & (vps - 1), y = (vy + ((c >> 1) & 1)) & (vps - 1),
              z = (vz + (c & 1)) & (vps - 1);
    float v = __int_as_float(0x7fc00000);
    if (s[c] >= 0) {
      const float2 q = dw[(size_t)s[c] * vpb + x + (y << sh) + (z << (2 * sh))];
      if (q.y > 1e-6f) v = q.x;
    }
    o[c] = v;
  }
  float4* dst = reinterpret_cast<float4*>(view + i * 8);
  dst[0] = make_float4(o[0], o[1], o[2], o[3]);
  dst[1] = make_float4(o[4], o[5], o[6], o[7]);
}

__global__ void deinterleave_kernel(const float2* __restrict__ dw, float* __restrict__ d,
                                    float* __restrict__ w, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 v = dw[i];
  d[i] = v.x;
  w[i] = v.y;
}

// Dense block index over the blocks' AABB (host-built: off the hot path, once per finished submap).
#define VGX_GRID_MAX_CELLS 4096
int vgx_submap_build_grid(vgx_ctx* c, VgxSubmap* s) {
  cudaFree(s->d_grid);
  cudaFree(s->d_grid16);
  s->d_grid = nullptr;
  s->d_grid16 = nullptr;
  s->grid_dim[0] = s->grid_dim[1] = s->grid_dim[2] = 0;
  const int n = s->n_blocks;
  if (n <= 0) return VGX_OK;
  std::vector<int32_t> idx(3 * (size_t)n);
  VGX_CUDA(c, cudaMemcpyAsync(idx.data(), s->d_block_idx, sizeof(int32_t) * 3 * n, cudaMemcpyDeviceToHost,
                              c->stream));
  VGX_CUDA(c, cudaStreamSynchronize(c->stream));
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], idx[3 * i + a]);
      hi[a] = std::max(hi[a], idx[3 * i + a]);
    }
  const long long cells = (long long)(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1);
  // sparse / huge: registration uses the hash (the kernel stages the grid as 16-bit slots)
  if (cells > VGX_GRID_MAX_CELLS || n >= 0xFFFF) return VGX_OK;
  std::vector<int32_t> grid((size_t)cells, -1);
  const int dx = hi[0] - lo[0] + 1, dy = hi[1] - lo[1] + 1;
  for (int i = 0; i < n; ++i)
    grid[(size_t)(idx[3 * i + 2] - lo[2]) * dy * dx + (size_t)(idx[3 * i + 1] - lo[1]) * dx +
         (idx[3 * i] - lo[0])] = i;
  VGX_CUDA(c, cudaMalloc(&s->d_grid, sizeof(int32_t) * cells));
  VGX_CUDA(c, cudaMemcpyAsync(s->d_grid, grid.data(), sizeof(int32_t) * cells, cudaMemcpyHostToDevice, c->stream));
  // 16-bit copy for the registration kernel's shared-memory grid (one bulk copy per tile)
  std::vector<uint16_t> g16(((size_t)cells + 7) & ~(size_t)7, 0xFFFF);
  for (long long k = 0; k < cells; ++k) g16[(size_t)k] = grid[(size_t)k] < 0 ? 0xFFFF : (uint16_t)grid[(size_t)k];
  VGX_CUDA(c, cudaMalloc(&s->d_grid16, sizeof(uint16_t) * g16.size()));
  VGX_CUDA(c, cudaMemcpyAsync(s->d_grid16, g16.data(), sizeof(uint16_t) * g16.size(), cudaMemcpyHostToDevice, c->stream));
  VGX_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int a = 0; a < 3; ++a) { s->grid_min[a] = lo[a]; s->grid_dim[a] = hi[a] - lo[a] + 1; }
  return VGX_OK;
}

int vgx_submap_build_view(vgx_ctx* c, VgxSubmap* s, const float2* bricks, float* view) {
  const size_t used = (size_t)s->n_blocks * s->vox_per_block;
  if (used == 0) return VGX_OK;
  int sh = 0;
  while ((1 << sh) < s->vps) sh++;
  build_view_kernel<<<(unsigned)((used + 255) / 256), 256, 0, c->stream>>>(bricks, s->d_block_idx, s->hash, view,
                                                                          s->n_blocks, s->vps, sh);
  c->launches++;
  VGX_CUDA(c, cudaGetLastError());
  return VGX_OK;
}

int vgx_submap_rebuild_hash(vgx_ctx* c, VgxSubmap* s) {
  const uint32_t tsize = s->hash.mask + 1;
  hash_clear_kernel<<<(tsize + 255) / 256, 256, 0, c->stream>>>(s->hash.entries, tsize);
  if (s->n_blocks > 0)
    hash_insert_kernel<<<(s->n_blocks + 127) / 128, 128, 0, c->stream>>>(s->hash, s->d_block_idx, s->n_blocks);
  c->launches += 2;
  VGX_CUDA(c, cudaGetLastError());
  return VGX_OK;
}

// ------------------------------------------------------------------ submap management
static uint32_t table_size_for(int cap) {
  uint32_t t = 64;
  while (t < (uint32_t)cap * 2u) t <<= 1;
  return t;
}

static int alloc_submap(vgx_ctx* c, uint32_t id, float voxel_size, int vps, int cap, bool with_view,
                        VgxSubmap** out) {
  if (!(voxel_size > 0) || vps < 2 || vps > 32 || (vps & (vps - 1)) != 0 || cap < 0)
    VGX_FAIL(c, VGX_ERR_INVALID, "invalid voxel_size / voxels_per_side / capacity");
  VGX_CUDA(c, cudaSetDevice(c->device));
  VgxSubmap* old = c->find(id);
  if (old) {
    VGX_CUDA(c, cudaStreamSynchronize(c->stream));
    free_submap(old);
    c->submaps.erase(id);
    vgx_graph_invalidate_registration(c);
  }
  VgxSubmap* s = new (std::nothrow) VgxSubmap();
  if (!s) return VGX_ERR_NOMEM;
  s->id = id;
  s->voxel_size = voxel_size;
  // voxblox Layer ctor: inverses by double division, stored as float
  s->voxel_size_inv = (float)(1.0 / (double)voxel_size);
  s->block_size = voxel_size * (float)vps;
  s->block_size_inv = (float)(1.0 / (double)s->block_size);
  s->vps = vps;
  s->vox_per_block = vps * vps * vps;
  s->cap_blocks = cap > 0 ? cap : 1;
  const uint32_t tsize = table_size_for(s->cap_blocks);
  s->hash.mask = tsize - 1;
  const size_t nvox = (size_t)s->cap_blocks * s->vox_per_block;
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaMalloc(&s->hash.entries, sizeof(VgxHashEntry) * tsize);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_block_idx, sizeof(int32_t) * 3 * s->cap_blocks);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_dw, sizeof(float2) * nvox);
  if (e == cudaSuccess && with_view) e = cudaMalloc(&s->d_view, sizeof(float) * 8 * nvox);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_counters, sizeof(int) * 4);
  if (e != cudaSuccess) {
    free_submap(s);
    c->set_error(std::string("submap allocation: ") + cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? VGX_ERR_NOMEM : VGX_ERR_CUDA;
  }
  hash_clear_kernel<<<(tsize + 255) / 256, 256, 0, c->stream>>>(s->hash.entries, tsize);
  c->launches++;
  VGX_CUDA(c, cudaMemsetAsync(s->d_counters, 0, sizeof(int) * 4, c->stream));
  c->submaps[id] = s;
  *out = s;
  return VGX_OK;
}

extern "C" int vgx_submap_upload(vgx_ctx* c, uint32_t id, float voxel_size, int vps, int n_blocks,
                                 const int32_t* block_idx, const float* distance,
                                 const float* weight) {
  if (!c) return VGX_ERR_INVALID;
  if (n_blocks < 0 || (n_blocks > 0 && (!block_idx || !distance || !weight)))
    VGX_FAIL(c, VGX_ERR_INVALID, "vgx_submap_upload: null input");
  VgxSubmap* s = nullptr;
  int rc = alloc_submap(c, id, voxel_size, vps, n_blocks, true, &s);
  if (rc != VGX_OK) return rc;
  s->n_blocks = n_blocks;
  s->finished = true;
  if (n_blocks == 0) return VGX_OK;
  const size_t nvox = (size_t)n_blocks * s->vox_per_block;
  rc = c->ensure_scratch(2 * nvox * sizeof(float));
  if (rc != VGX_OK) return rc;
  float* tmp_d = (float*)c->d_scratch;
  float* tmp_w = tmp_d + nvox;
  VGX_CUDA(c, cudaMemcpyAsync(tmp_d, distance, nvox * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  VGX_CUDA(c, cudaMemcpyAsync(tmp_w, weight, nvox * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  VGX_CUDA(c, cudaMemcpyAsync(s->d_block_idx, block_idx, sizeof(int32_t) * 3 * n_blocks,
                              cudaMemcpyHostToDevice, c->stream));
  VGX_CUDA(c, cudaMemcpyAsync(s->d_counters, &s->n_blocks, sizeof(int), cudaMemcpyHostToDevice,
                              c->stream));
  hash_insert_kernel<<<(n_blocks + 127) / 128, 128, 0, c->stream>>>(s->hash, s->d_block_idx, n_blocks);
  interleave_kernel<<<(unsigned)((nvox + 255) / 256), 256, 0, c->stream>>>(tmp_d, tmp_w, s->d_dw, nvox);
  {
    int sh = 0;
    while ((1 << sh) < vps) sh++;
    build_view_kernel<<<(unsigned)((nvox + 255) / 256), 256, 0, c->stream>>>(
        s->d_dw, s->d_block_idx, s->hash, s->d_view, n_blocks, vps, sh);
  }
  c->launches += 3;
  VGX_CUDA(c, cudaGetLastError());
  VGX_CUDA(c, cudaStreamSynchronize(c->stream));  // host buffers are the caller's
  return vgx_submap_build_grid(c, s);
}

extern "C" int vgx_submap_peek_device(vgx_ctx* c, uint32_t id, const int32_t** d_block_idx,
                                      const float** d_dw, int* n_blocks) {
  if (!c) return VGX_ERR_INVALID;
  VgxSubmap* s = c->find(id);
  if (!s) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_submap_peek_device: unknown submap");
  VGX_CUDA(c, cudaSetDevice(c->device));
  VGX_CUDA(c, cudaStreamSynchronize(c->stream));   // the caller reads the bricks on its own stream
  if (d_block_idx) *d_block_idx = s->d_block_idx;
  if (d_dw) *d_dw = reinterpret_cast<const float*>(s->d_dw);
  if (n_blocks) *n_blocks = s->n_blocks;
  return VGX_OK;
}

extern "C" int vgx_submap_upload_device(vgx_ctx* c, uint32_t id, float voxel_size, int vps, int n_blocks,
                                        const int32_t* d_block_idx, const float* d_dw) {
  if (!c) return VGX_ERR_INVALID;
  if (n_blocks < 0 || (n_blocks > 0 && (!d_block_idx || !d_dw)))
    VGX_FAIL(c, VGX_ERR_INVALID, "vgx_submap_upload_device: null input");
  VgxSubmap* s = nullptr;
  int rc = alloc_submap(c, id, voxel_size, vps, n_blocks, true, &s);
  if (rc != VGX_OK) return rc;
  s->n_blocks = n_blocks;
  s->finished = true;
  if (n_blocks == 0) return VGX_OK;
  const size_t nvox = (size_t)n_blocks * s->vox_per_block;
  VGX_CUDA(c, cudaMemcpyAsync(s->d_block_idx, d_block_idx, sizeof(int32_t) * 3 * n_blocks, cudaMemcpyDeviceToDevice, c->stream));
  VGX_CUDA(c, cudaMemcpyAsync(s->d_dw, d_dw, sizeof(float2) * nvox, cudaMemcpyDeviceToDevice, c->stream));
  VGX_CUDA(c, cudaMemcpyAsync(s->d_counters, &s->n_blocks, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  hash_insert_kernel<<<(n_blocks + 127) / 128, 128, 0, c->stream>>>(s->hash, s->d_block_idx, n_blocks);
  c->launches++;
  rc = vgx_submap_build_view(c, s, s->d_dw, s->d_view);
  if (rc != VGX_OK) return rc;
  VGX_CUDA(c, cudaStreamSynchronize(c->stream));   // the source buffers are the caller's
  return vgx_submap_build_grid(c, s);
}

extern "C" int vgx_submap_create(vgx_ctx* c, uint32_t id, float voxel_size, int vps, int capacity) {
  if (!c) return VGX_ERR_INVALID;
  if (capacity <= 0) VGX_FAIL(c, VGX_ERR_INVALID, "vgx_submap_create: capacity must be > 0");
  VgxSubmap* s = nullptr;
  int rc = alloc_submap(c, id, voxel_size, vps, capacity, false, &s);
  if (rc != VGX_OK) return rc;
  // new voxblox blocks are zero-initialised (distance 0, weight 0)
  VGX_CUDA(c, cudaMemsetAsync(s->d_dw, 0, sizeof(float2) * (size_t)capacity * s->vox_per_block,
                              c->stream));
  s->finished = false;
  return VGX_OK;
}

extern "C" int vgx_submap_finish(vgx_ctx* c, uint32_t id) {
  if (!c) return VGX_ERR_INVALID;
  VgxSubmap* s = c->find(id);
  if (!s) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_submap_finish: unknown submap");
  VGX_CUDA(c, cudaSetDevice(c->device));
  // the layer is frozen from here on: the view only needs the blocks that exist, not the
  // integration capacity (128 KiB per brick)
  const size_t used = (size_t)s->n_blocks * s->vox_per_block;
  if (!s->d_view) VGX_CUDA(c, cudaMalloc(&s->d_view, sizeof(float) * 8 * (used > 0 ? used : 1)));
  if (used > 0) {
    int sh = 0;
    while ((1 << sh) < s->vps) sh++;
    build_view_kernel<<<(unsigned)((used + 255) / 256), 256, 0, c->stream>>>(
        s->d_dw, s->d_block_idx, s->hash, s->d_view, s->n_blocks, s->vps, sh);
    c->launches++;
    VGX_CUDA(c, cudaGetLastError());
  }
  s->finished = true;
  vgx_graph_invalidate_registration(c);
  return vgx_submap_build_grid(c, s);
}

extern "C" int vgx_submap_free(vgx_ctx* c, uint32_t id) {
  if (!c) return VGX_ERR_INVALID;
  VgxSubmap* s = c->find(id);
  if (!s) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_submap_free: unknown submap");
  VGX_CUDA(c, cudaSetDevice(c->device));
  VGX_CUDA(c, cudaStreamSynchronize(c->stream));
  free_submap(s);
  c->submaps.erase(id);
  vgx_graph_invalidate_registration(c);
  return VGX_OK;
}

extern "C" int vgx_submap_block_count(vgx_ctx* c, uint32_t id, int* n) {
  if (!c || !n) return VGX_ERR_INVALID;
  VgxSubmap* s = c->find(id);
  if (!s) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_submap_block_count: unknown submap");
  *n = s->n_blocks;
  return VGX_OK;
}

extern "C" int vgx_submap_info(vgx_ctx* c, uint32_t id, float* voxel_size, int* vps, int* n_blocks,
                               int* finished) {
  if (!c) return VGX_ERR_INVALID;
  VgxSubmap* s = c->find(id);
  if (!s) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_submap_info: unknown submap");
  if (voxel_size) *voxel_size = s->voxel_size;
  if (vps) *vps = s->vps;
  if (n_blocks) *n_blocks = s->n_blocks;
  if (finished) *finished = s->finished ? 1 : 0;
  return VGX_OK;
}

extern "C" int vgx_submap_download(vgx_ctx* c, uint32_t id, int max_blocks, int32_t* block_idx,
                                   float* distance, float* weight, int* n_out) {
  if (!c) return VGX_ERR_INVALID;
  VgxSubmap* s = c->find(id);
  if (!s) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_submap_download: unknown submap");
  VGX_CUDA(c, cudaSetDevice(c->device));
  const int n = s->n_blocks;
  if (n_out) *n_out = n;
  if (n > max_blocks) VGX_FAIL(c, VGX_ERR_CAPACITY, "vgx_submap_download: max_blocks too small");
  if (n == 0) return VGX_OK;
  const size_t nvox = (size_t)n * s->vox_per_block;
  if (block_idx)
    VGX_CUDA(c, cudaMemcpyAsync(block_idx, s->d_block_idx, sizeof(int32_t) * 3 * n,
                                cudaMemcpyDeviceToHost, c->stream));
  if (distance || weight) {
    int rc = c->ensure_scratch(2 * nvox * sizeof(float));
    if (rc != VGX_OK) return rc;
    float* tmp_d = (float*)c->d_scratch;
    float* tmp_w = tmp_d + nvox;
    deinterleave_kernel<<<(unsigned)((nvox + 255) / 256), 256, 0, c->stream>>>(s->d_dw, tmp_d, tmp_w, nvox);
    c->launches++;
    VGX_CUDA(c, cudaGetLastError());
    if (distance)
      VGX_CUDA(c, cudaMemcpyAsync(distance, tmp_d, nvox * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (weight)
      VGX_CUDA(c, cudaMemcpyAsync(weight, tmp_w, nvox * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  }
  VGX_CUDA(c, cudaStreamSynchronize(c->stream));
  return VGX_OK;
}

extern "C" int vgx_submap_upload_points(vgx_ctx* c, uint32_t id, int type, int n, const float* xyz,
                                        const float* distance, const float* weight) {
  if (!c) return VGX_ERR_INVALID;
  if (type < 0 || type > 1 || n < 0 || (n > 0 && (!xyz || !distance || !weight)))
    VGX_FAIL(c, VGX_ERR_INVALID, "vgx_submap_upload_points: invalid argument");
  VgxSubmap* s = c->find(id);
  if (!s) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_submap_upload_points: unknown submap");
  VGX_CUDA(c, cudaSetDevice(c->device));
  VGX_CUDA(c, cudaStreamSynchronize(c->stream));
  VgxPoints& p = s->points[type];
  free_points(p);
  vgx_graph_invalidate_registration(c);
  if (n == 0) return VGX_OK;
  // unit-major AoSoA staging on the host (pinned): unit u = {x[32], y[32], z[32], d[32], w[32]}
  const size_t units = ((size_t)n + VGX_PT_UNIT - 1) / VGX_PT_UNIT;
  const size_t floats = units * VGX_PT_UNIT_FLOATS;
  int rc = c->ensure_pinned(floats * sizeof(float));
  if (rc != VGX_OK) return rc;
  float* h = (float*)c->h_pinned;
  memset(h, 0, floats * sizeof(float));
  double sum_w = 0;
  std::vector<double> cumulative((size_t)n);
  for (int i = 0; i < n; ++i) {
    float* u = h + vgx_pt_index((size_t)i, 0);
    u[0] = xyz[3 * (size_t)i];
    u[32] = xyz[3 * (size_t)i + 1];
    u[64] = xyz[3 * (size_t)i + 2];
    u[96] = distance[i];
    u[128] = weight[i];
    sum_w += (double)weight[i];  // cpp:124, in point order
    cumulative[i] = sum_w;       // WeightedSampler::addItem (weighted_sampler_inl.h:6-17)
  }
  float* d = nullptr;
  VGX_CUDA(c, cudaMalloc(&d, floats * sizeof(float)));
  cudaError_t e = cudaMemcpyAsync(d, h, floats * sizeof(float), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    cudaFree(d);
    c->set_error(std::string("points upload: ") + cudaGetErrorString(e));
    return VGX_ERR_CUDA;
  }
  p.n = n;
  p.data = d;
  p.sum_w = sum_w;
  p.cumulative_w.swap(cumulative);
  return VGX_OK;
}

// ------------------------------------------------------------------ WeightedSampler
// weighted_sampler_inl.h:19-28: r = uniform_real_distribution<double>(0,1)(mt19937),
// index = upper_bound(cumulative, r * cumulative.back()).  std::mt19937's default seed (5489)
// and uniform_real_distribution are what the reference instantiates (weighted_sampler.h:34-36).
// fl(r * back) can round up to back itself, where upper_bound returns end(): the reference then
// indexes one past the last item; clamped here.
void vgx_points_draw(VgxPoints& p, int count, int32_t* idx) {
  std::uniform_real_distribution<double> uniform(0.0, 1.0);
  const std::vector<double>& cw = p.cumulative_w;
  for (int i = 0; i < count; ++i) {
    const double r = uniform(p.rng);
    const double t = r * cw.back();
    size_t k = (size_t)(std::upper_bound(cw.begin(), cw.end(), t) - cw.begin());
    if (k >= cw.size()) k = cw.size() - 1;
    idx[i] = (int32_t)k;
  }
}

extern "C" int vgx_submap_draw_samples(vgx_ctx* c, uint32_t id, int type, int n, int32_t* indices) {
  if (!c) return VGX_ERR_INVALID;
  if (type < 0 || type > 1 || n < 0 || (n > 0 && !indices))
    VGX_FAIL(c, VGX_ERR_INVALID, "vgx_submap_draw_samples: invalid argument");
  VgxSubmap* s = c->find(id);
  if (!s) VGX_FAIL(c, VGX_ERR_NOT_FOUND, "vgx_submap_draw_samples: unknown submap");
  VgxPoints& p = s->points[type];
  if (p.n == 0) VGX_FAIL(c, VGX_ERR_INVALID, "vgx_submap_draw_samples: the submap has no registration points");
  vgx_points_draw(p, n, indices);
  return VGX_OK;
}
