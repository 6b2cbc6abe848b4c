// Stand-alone "tanh + neighbour max-pool" of the encoders' layer loop, for callers that chain GraphConvolution layers themselves:
//     T = tanh(y);   P[b,i,c] = max_j adj[b,i,j] * T[b,j,c]   (first index wins ties)        forward
//     dT[b,j,c] = sum_{i : arg[b,i,c] == j} adj[b,i,j] * dP[b,i,c];   dy = dT * (1 - T^2)        backward
// Reference: openchem/modules/encoders/gcn_encoder.py:43-50 and edge_attention_encoder.py:54-61 (x.repeat -> view(B,N,N,D) ->
// * adj.expand -> max(dim=2), after torch.tanh). The fused GraphCNNEncoder path does this inside the tile kernel; this kernel
// serves GraphEdgeAttentionEncoder, whose pool adjacency (edge_attr[..., 0] > 0) differs from the soft adjacencies its
// GraphConvolution layers multiply with. Exact fp32 (tanhf), dense scan in the reference's order, any float adjacency, any strides.
//
// One CTA per (molecule, block of 128 channels), one thread per channel: a thread owns its channel column in both directions, so
// the backward scatter needs no atomics and every sum has a fixed order (bit-reproducible).
#pragma once
#include "ocgcn_common.cuh"

namespace ocgcn {

constexpr int kPoolCols = 128;
static inline size_t pool_smem_bytes(int N, bool bwd) { return (size_t)N * kPoolCols * sizeof(float) * (bwd ? 2 : 1); }

struct PoolArgs {
  int B, N, D;
  const float* y;             // (B,N,D) contiguous
  const float* adj;           // (B,N,N), element strides as0 (molecule), as1 (row i), as2 (column j)
  long long as0, as1, as2;
  float* pooled;              // fwd out (B,N,D)
  unsigned char* arg;         // fwd out / bwd in (B,N,D)
  const float* grad_pooled;   // bwd in (B,N,D)
  float* grad_y;              // bwd out (B,N,D)
};

__global__ void __launch_bounds__(kPoolCols) pool_fwd_kernel(const PoolArgs a) {
  extern __shared__ __align__(16) float psm[];
  pdl_wait();
  const int b = blockIdx.y, c = blockIdx.x * kPoolCols + threadIdx.x, N = a.N, D = a.D;
  const bool ok = c < D;
  const float* yb = a.y + (size_t)b * N * D;
  for (int j = 0; j < N; ++j) psm[j * kPoolCols + threadIdx.x] = ok ? tanhf(__ldg(&yb[(size_t)j * D + c])) : 0.f;
  // (a thread reads back only what it wrote: no barrier needed)
  const float* ab = a.adj + (long long)b * a.as0;
  for (int i = 0; i < N; ++i) {
    const float* arow = ab + (long long)i * a.as1;
    float best = __ldg(&arow[0]) * psm[threadIdx.x];
    int arg = 0;
    for (int j = 1; j < N; ++j) {
      const float cand = __ldg(&arow[(long long)j * a.as2]) * psm[j * kPoolCols + threadIdx.x];
      if (cand > best) { best = cand; arg = j; }
    }
    if (ok) {
      a.pooled[((size_t)b * N + i) * D + c] = best;
      a.arg[((size_t)b * N + i) * D + c] = (unsigned char)arg;
    }
  }
}

__global__ void __launch_bounds__(kPoolCols) pool_bwd_kernel(const PoolArgs a) {
  extern __shared__ __align__(16) float psm[];
  pdl_wait();
  const int b = blockIdx.y, c = blockIdx.x * kPoolCols + threadIdx.x, N = a.N, D = a.D;
  if (c >= D) return;
  float* sT = psm;                          // [N][128] tanh(y)
  float* sG = psm + (size_t)N * kPoolCols;  // [N][128] dT
  const float* yb = a.y + (size_t)b * N * D;
  for (int j = 0; j < N; ++j) {
    sT[j * kPoolCols + threadIdx.x] = tanhf(__ldg(&yb[(size_t)j * D + c]));
    sG[j * kPoolCols + threadIdx.x] = 0.f;
  }
  const float* ab = a.adj + (long long)b * a.as0;
  for (int i = 0; i < N; ++i) {   // ascending i: the order in which autograd's sum over the repeated axis adds the contributions
    const size_t e = ((size_t)b * N + i) * D + c;
    const int j = a.arg[e];
    sG[j * kPoolCols + threadIdx.x] += __ldg(&ab[(long long)i * a.as1 + (long long)j * a.as2]) * __ldg(&a.grad_pooled[e]);
  }
  for (int j = 0; j < N; ++j) {
    const float t = sT[j * kPoolCols + threadIdx.x];
    a.grad_y[((size_t)b * N + j) * D + c] = sG[j * kPoolCols + threadIdx.x] * (1.f - t * t);
  }
}

}  // namespace ocgcn
