// Compact device batch format: bit-packed adjacency and bit-packed binary node features.
//
// The reference uploads dense fp32 `adj_matrix (B,N,N)` and `node_feature_matrix (B,N,F0)` synchronously every iteration
// (Graph2Label.cast_inputs, openchem/models/Graph2Label.py:41-57): 4N^2 + 4N*F0 bytes per molecule. Both are 0/1 matrices
// for every featuriser the reference ships (one-hot groups, utils/graph.py:57-63; bonds + self loops, :80-84), so the same
// information fits in N*ceil(N/32) + N*ceil(F0/32) 32-bit words (1 KiB instead of 32 KiB per molecule at N = F0 = 64).
//
//   bits[(mol * N + row) * W + (col >> 5)]  bit (col & 31)  <=>  dense[mol][row][col] == 1.0f          W = ceil(C / 32)
//
// expand_bits_kernel writes the dense fp32 tensors the encoder entry points take (bit-exact: only 0.0f and 1.0f occur),
// optionally gathering molecules by index from a compact data set resident in HBM; pack_bits_kernel is the inverse and
// reports non-binary input instead of silently rounding it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ocgcn {

// one thread per group of four consecutive columns; rows = B*N output rows of C columns
// `index` entries outside [0, n_src) are never dereferenced: the molecule is written as zeros and *oob (integer flag, order-independent)
// is raised, like pack_bits_kernel's non-binary flag
__global__ void __launch_bounds__(256) expand_bits_kernel(const uint32_t* __restrict__ bits, const int* __restrict__ index, int n_src,
                                                          int* __restrict__ oob, int N, int C, int W, long long rows,
                                                          float* __restrict__ out) {
  const int C4 = (C + 3) >> 2;
  const long long total = rows * C4;
  const bool vec = (C & 3) == 0;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / C4;
    const int c = (int)(idx - row * C4) << 2;
    long long srow = row;
    bool ok = true;
    if (index) {
      const long long b = row / N;
      const int src = __ldg(&index[b]);
      ok = src >= 0 && src < n_src;
      if (!ok && c == 0 && row == b * N) atomicOr(oob, 1);
      srow = (long long)src * N + (row - b * N);
    }
    const uint32_t nib = ok ? (__ldg(&bits[srow * W + (c >> 5)]) >> (c & 31)) & 0xfu : 0u;   // c is a multiple of 4: one word holds all four
    const float v0 = (nib & 1u) ? 1.f : 0.f, v1 = (nib & 2u) ? 1.f : 0.f, v2 = (nib & 4u) ? 1.f : 0.f, v3 = (nib & 8u) ? 1.f : 0.f;
    float* dst = out + row * C + c;
    if (vec) {
      __stcs(reinterpret_cast<float4*>(dst), make_float4(v0, v1, v2, v3));
    } else {
      dst[0] = v0;
      if (c + 1 < C) dst[1] = v1;
      if (c + 2 < C) dst[2] = v2;
      if (c + 3 < C) dst[3] = v3;
    }
  }
}

// one warp per output word: lane = column inside the word; strided dense input (element strides s0 molecule, s1 row, s2 column)
__global__ void __launch_bounds__(256) pack_bits_kernel(const float* __restrict__ dense, long long s0, long long s1, long long s2, int N,
                                                        int C, int W, long long rows, uint32_t* __restrict__ bits,
                                                        int* __restrict__ nonbinary) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  int bad = 0;
  for (long long wd = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wd < rows * W; wd += warps) {
    const long long row = wd / W;
    const int w = (int)(wd - row * W), c = w * 32 + lane;
    const long long b = row / N, i = row - b * N;
    const float v = c < C ? __ldg(&dense[b * s0 + i * s1 + (long long)c * s2]) : 0.f;
    bad |= (v != 0.f) & (v != 1.f);
    const uint32_t word = __ballot_sync(0xffffffffu, v != 0.f);
    if (lane == 0) bits[wd] = word;
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(nonbinary, 1);   // integer flag: order-independent
}

}  // namespace ocgcn
