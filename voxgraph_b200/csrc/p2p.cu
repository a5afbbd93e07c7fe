// One-shot all-gather-reduce of the packed normal equations over NVLink peer memory (CUDA IPC).
// See vgx_comm_p2p_export / vgx_comm_p2p_import in include/voxgraph_b200.h.
//
// Per evaluation (epoch e, parity e & 1):
//   assemble_kernel      writes this rank's partial into its own exported buffer[parity]
//   (its last CTA)       __threadfence_system(), then stores e into flag[rank] of EVERY rank's region
//   p2p_gather_kernel    waits until all flags in its own region reach e, then sums the partials of
//                        ranks 0..n-1 in rank order straight out of peer memory (ld.volatile, no L1)
// Two buffers suffice: a rank can only signal epoch e+1 after its own gather of e has finished
// (stream order), and nobody passes the gather of e+1 before everybody signalled e+1.
#include <stdlib.h>
#include <string.h>

#include "vgx_internal.h"

#define P2P_FLAG_BYTES 256

struct P2PPeers {
  double* buf[8];                 // buffer[parity] of every rank
  unsigned long long* flags[8];   // flag array of every rank
  int nranks, rank;
};

__global__ void __launch_bounds__(256)
p2p_gather_kernel(P2PPeers P, unsigned long long epoch, double* __restrict__ out, size_t count,
                  int* __restrict__ timeout_flag) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    const volatile unsigned long long* mine = P.flags[P.rank];
    const long long t0 = clock64();
    int ok = 1;
    for (int r = 0; r < P.nranks; ++r) {
      while (mine[r] < epoch) {
        if (clock64() - t0 > 4000000000ll) { ok = 0; break; }  // ~2 s: a peer is gone
      }
      if (!ok) break;
    }
    __threadfence_system();
    s_ok = ok;
    if (!ok) *timeout_flag = 1;
  }
  __syncthreads();
  if (!s_ok) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < P.nranks; ++r) s += *((const volatile double*)(P.buf[r] + i));
    out[i] = s;
  }
}

void vgx_p2p_free(vgx_ctx* c) {
  if (!c->p2p_base) return;
  cudaStreamSynchronize(c->stream);
  for (int r = 0; r < 8; ++r)
    if (c->p2p_peer[r] && c->p2p_peer[r] != c->p2p_base) cudaIpcCloseMemHandle(c->p2p_peer[r]);
  cudaFree(c->p2p_base);
  c->p2p_base = nullptr;
  memset(c->p2p_peer, 0, sizeof(c->p2p_peer));
  c->p2p_ready = false;
  c->p2p_cap = 0;
}

extern "C" int vgx_comm_p2p_export(vgx_ctx* c, uint64_t capacity_doubles, uint8_t handle[64]) {
  if (!c || !handle || capacity_doubles == 0) return VGX_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  VGX_CUDA(c, cudaSetDevice(c->device));
  vgx_p2p_free(c);
  const size_t cap = ((size_t)capacity_doubles + 31) & ~(size_t)31;
  const size_t bytes = P2P_FLAG_BYTES + 2 * cap * sizeof(double);
  VGX_CUDA(c, cudaMalloc(&c->p2p_base, bytes));
  VGX_CUDA(c, cudaMemset(c->p2p_base, 0, bytes));
  cudaIpcMemHandle_t h;
  VGX_CUDA(c, cudaIpcGetMemHandle(&h, c->p2p_base));
  memcpy(handle, &h, 64);
  c->p2p_cap = cap;
  c->p2p_epoch = 0;
  return VGX_OK;
}

extern "C" int vgx_comm_p2p_import(vgx_ctx* c, int nranks, int rank, const uint8_t* handles) {
  if (!c || !handles || nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks) return VGX_ERR_INVALID;
  if (!c->p2p_base) VGX_FAIL(c, VGX_ERR_INVALID, "vgx_comm_p2p_import: call vgx_comm_p2p_export first");
  if (c->nccl_comm && (c->nranks != nranks || c->rank != rank))
    VGX_FAIL(c, VGX_ERR_INVALID, "vgx_comm_p2p_import: rank layout differs from vgx_comm_init");
  VGX_CUDA(c, cudaSetDevice(c->device));
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) { c->p2p_peer[r] = c->p2p_base; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)r, 64);
    void* p = nullptr;
    VGX_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->p2p_peer[r] = p;
  }
  c->nranks = nranks;
  c->rank = rank;
  c->p2p_ready = true;
  {
    // the one-launch path is opt-in until it has been validated on every rank count
    const char* f = getenv("VGX_P2P_FUSED");
    c->p2p_fused = f && f[0] == '1';
  }
  vgx_graph_invalidate_registration(c);
  return VGX_OK;
}

int vgx_p2p_check(vgx_ctx* c) {
  if (!c->p2p_ready) return VGX_OK;
  int t = 0;
  VGX_CUDA(c, cudaMemcpyAsync(&t, (char*)c->p2p_base + 128, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  VGX_CUDA(c, cudaStreamSynchronize(c->stream));
  if (t) VGX_FAIL(c, VGX_ERR_NCCL, "peer exchange timed out waiting for a rank (NVLink all-gather-reduce)");
  return VGX_OK;
}

int vgx_p2p_begin(vgx_ctx* c, size_t count, double** send_buf, VgxP2PSignal* sig) {
  if (!c->p2p_ready) VGX_FAIL(c, VGX_ERR_INVALID, "peer exchange not initialised");
  if (count > c->p2p_cap) VGX_FAIL(c, VGX_ERR_CAPACITY, "packed normal equations exceed the exported peer buffer");
  const unsigned long long e = ++c->p2p_epoch;
  *send_buf = (double*)((char*)c->p2p_base + P2P_FLAG_BYTES) + (e & 1) * c->p2p_cap;
  memset(sig, 0, sizeof(*sig));
  sig->epoch = e;
  sig->nranks = c->nranks;
  sig->rank = c->rank;
  sig->counter = (int*)((char*)c->p2p_base + 192);  // local word of the flag page
  for (int r = 0; r < c->nranks; ++r) sig->flags[r] = (unsigned long long*)c->p2p_peer[r];
  return VGX_OK;
}

void vgx_p2p_gather_sources(vgx_ctx* c, const void* bufs[8], int** timeout_flag) {
  const unsigned long long e = c->p2p_epoch;
  for (int r = 0; r < 8; ++r)
    bufs[r] = r < c->nranks ? (const void*)((double*)((char*)c->p2p_peer[r] + P2P_FLAG_BYTES) + (e & 1) * c->p2p_cap)
                            : nullptr;
  *timeout_flag = (int*)((char*)c->p2p_base + 128);
}

int vgx_p2p_gather(vgx_ctx* c, double* d_out, size_t count) {
  const unsigned long long e = c->p2p_epoch;
  P2PPeers P;
  memset(&P, 0, sizeof(P));
  P.nranks = c->nranks;
  P.rank = c->rank;
  for (int r = 0; r < c->nranks; ++r) {
    P.flags[r] = (unsigned long long*)c->p2p_peer[r];
    P.buf[r] = (double*)((char*)c->p2p_peer[r] + P2P_FLAG_BYTES) + (e & 1) * c->p2p_cap;
  }
  int* d_timeout = (int*)((char*)c->p2p_base + 128);  // local word of the flag page, peers never touch it
  {
    VgxLaunchScope s(c, 5);
    const int blocks = (int)((count + 1023) / 1024) > 0 ? (int)((count + 1023) / 1024) : 1;
    p2p_gather_kernel<<<blocks < 64 ? blocks : 64, 256, 0, c->stream>>>(P, e, d_out, count, d_timeout);
  }
  VGX_CUDA(c, cudaGetLastError());
  return VGX_OK;
}

extern "C" int vgx_comm_suspend(vgx_ctx* c, int on) {
  if (!c) return VGX_ERR_INVALID;
  if (on) {
    if (c->saved_nranks == 0) {
      c->saved_nranks = c->nranks;
      c->saved_rank = c->rank;
      c->nranks = 1;
      c->rank = 0;
      vgx_graph_invalidate_registration(c);
    }
  } else if (c->saved_nranks != 0) {
    c->nranks = c->saved_nranks;
    c->rank = c->saved_rank;
    c->saved_nranks = 0;
    c->saved_rank = 0;
    vgx_graph_invalidate_registration(c);
  }
  return VGX_OK;
}
