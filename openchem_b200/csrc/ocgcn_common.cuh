// Shared definitions for the ocgcn kernels (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ocgcn.h"

namespace ocgcn {

constexpr int kThreads = 256;    // the exact-fp32 tile kernels and the streaming kernels run 8 warps
constexpr int kThreadsMM = 512;  // tensor-core tile kernels: 16 warps (accumulators live in TMEM, <= 64 registers per thread)
constexpr int kWarps = kThreads / 32;
constexpr int KC = 64;          // channel chunk: two adjacent channels per lane
constexpr int KCP = KC + 4;     // padded chunk row (floats): float4-aligned rows, conflict-free float2 per lane
constexpr int kMaxNbr = 13;     // neighbours held inline in a 16-byte row record (more -> bit-mask fallback)

typedef unsigned long long u64;

// Programmatic dependent launch (PDL). The kernels of one forward / backward call are launched back to back with the
// programmatic-stream-serialization attribute (ocgcn_api.cu: launch_chain): a kernel may become resident while its predecessor
// drains, runs its prologue (shared-memory carve-up, mbarrier init, TMEM allocation), and blocks in pdl_wait() until the
// predecessor grid has COMPLETED and its memory is visible - so pdl_wait() comes before the first global-memory access.
// pdl_trigger() marks the point after which a CTA no longer delays the dependent's launch (all CTAs triggered or exited).
// Both are no-ops when the kernel was launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- producer / epilogue kinds of the tile kernel ------------------------------------------------------------
enum Prod : int {
  P_X = 0,    // first layer: H = x (strided load);            S = A.H
  P_MID = 1,  // H = pool(tanh(scale*Zprev+shift)) (+arg-max);  S = A.H
  P_RO = 2,   // H = pool(tanh(scale*Zprev+shift)) (+arg-max);  GEMM operand = H   (readout)
  P_DZ = 3,   // dZ = g*(dY - c1 - Yhat*c2)                     (BatchNorm backward, second half)
  P_DU = 4    // dU = dO*(1-O^2)*(1-D^2)                        (readout backward)
};
enum Epi : int {
  E_ZSTATS = 0,   // Z = acc + bias -> global, per-tile Welford partials (count, mean, M2) per channel
  E_READOUT = 1,  // D = tanh(acc + bias) -> global, O = tanh(sum_i D) -> global
  E_STORE = 2,    // out = acc
  E_ATG = 3,      // out = A^T . acc  (per molecule)
  E_DY = 4,       // acc is dH of the previous layer: pool scatter through its arg-max + tanh' -> out = dY_prev, partial sums
  E_ATG_DY = 5    // dH = A^T . acc, then as E_DY  (fuses the first half of the previous layer's backward)
};

struct TileArgs {
  int B, N, NW, MT;  // molecules, atoms, mask words per row, molecules per tile
  int K, NC;         // GEMM reduction width (producer channels) and output columns
  // adjacency: dense (general path) + packed
  const float* adj;
  long long as0, as1, as2;
  const u64* rowmask;       // [B*N][NW] bit j of row i  <=> adj[b,i,j] != 0
  const u64* colmask;       // [B*N][NW] bit i of col j  <=> adj[b,i,j] != 0
  const uint4* nbr;         // [B*N] row records: see NbrRec
  const uint4* cnbr;        // [B*N] column records (neighbours of j in A^T)
  const unsigned* molflag;  // [B] 1 = every entry of adj[b] is exactly 0 or 1
  // producer inputs
  const float* in0;  // P_X: x | P_MID/P_RO: Zprev | P_DZ: dY | P_DU: D
  long long xs0, xs1, xs2;
  const uint32_t* x_bits;  // P_X, compact batch: bit (c & 31) of word x_bits[(mol*N + i)*xw + (c >> 5)] is x[mol,i,c] (in0 is null)
  int xw;                  // words per row of x_bits = ceil(F0/32)
  const float* in1;   // P_DZ: Z | P_DU: dO
  const float* in2;   // P_DU: O
  const float* vec0;  // P_MID/P_RO: scale | P_DZ: g = gamma*invstd
  const float* vec1;  // P_MID/P_RO: shift | P_DZ: c1 = dbeta/n
  const float* vec2;  // P_DZ: c2 = dgamma/n
  const float* vec3;  // P_DZ: mean
  const float* vec4;  // P_DZ: invstd
  // producer outputs (written on the first column pass only)
  unsigned char* arg_out;  // P_MID/P_RO  [rows][K]
  float* prod_out;         // P_X/P_MID: S | P_RO: H_L | P_DZ: dZ | P_DU: dU   [rows][K]
  float* prod_partials;    // P_DZ/P_DU: per-tile column sums [tile][K]
  // GEMM right operand [K][NC] row-major
  const float* Bm;
  const float* bias;  // NC or null
  // epilogue outputs
  float* out0;      // Z | D | dH | dHprev  [rows][NC]
  float* out1;      // E_READOUT: O [B][NC]
  float* partials;  // E_ZSTATS: [tile][2][NC] (mean, M2) | E_DY / E_ATG_DY: [tile][2][NC] (sum dY, sum dY*Yhat)
  int skip_gemm;    // producers only (P_DZ of the first layer when dX is not needed)
  // E_DY / E_ATG_DY: saved state of the PREVIOUS layer (the one whose pooled output this tile's dH is the gradient of)
  const unsigned char* arg_in;  // [rows][NC] pool arg-max
  const float* zprev;           // [rows][NC] pre-BatchNorm activations
  const float* pv_scale;        // [NC] gamma*invstd
  const float* pv_shift;        // [NC] beta - mean*scale
  const float* pv_mean;         // [NC]
  const float* pv_invstd;       // [NC]
  // tensor-core engine (tile_kernel<.., MM = 1>) only
  const void* wimg;  // weight operand images: pieces [K/64][ceil(NC/128)][hi(,lo)] of [8 k-chunks][128 n][8] bf16 (ocgcn_dw.cuh)
  int ncb;           // ceil(NC/128)
  void* prod_img;    // producer output as bf16 tile images [tile][img_chunks][128 rows][8] (operand of the dW GEMM) or null
  int img_chunks;    // 16-byte channel chunks per tile image = 8 * ceil(K/64)
  int ntiles;        // tiles of the batch (the MM kernels are persistent: a CTA loops over tiles blockIdx.x, + gridDim.x, ...)
  // fused weight gradient (P_DZ / P_DU of the tensor-core engine, widths <= 128): dW[m][k] = sum_rows saved[row][m] * produced[row][k]
  const void* dw_img;   // saved operand tile images [tile][dw_chunks][128][8] bf16 (S_l, or H_L for the readout), or null
  int dw_chunks;        // 16-byte channel chunks of that image per tile (<= 16)
  int dw_M;             // true row extent of the gradient (F_in, or F_L for the readout)
  float* dw_partial;    // [gridDim.x][dw_M][K] per-CTA partial gradients
  // TMA staging of the forward producers' input (P_MID / P_RO of the tensor-core engine): the 128-row x 64-channel fp32 chunk of
  // Zprev arrives through the kernel's tensor map (cp.async.bulk.tensor) instead of per-thread loads
  int use_tma;
  int reverse;   // tiles are visited last to first (see tile_kernel)
};

// x[mol, i, c] and x[mol, i, c + 1] (c even) of the first layer's input: from the dense fp32 tensor, or from the bit words of a
// compact batch (compact.py; the two bits share a word). `second` = channel c + 1 exists.
__device__ __forceinline__ float2 load_x_pair(const float* __restrict__ x, long long xs0, long long xs1, long long xs2,
                                              const uint32_t* __restrict__ x_bits, int xw, int N, long long mol, int i, int c, bool second) {
  if (x_bits) {
    const uint32_t w = __ldg(&x_bits[((size_t)mol * N + i) * xw + (c >> 5)]) >> (c & 31);
    return make_float2((float)(w & 1u), second ? (float)((w >> 1) & 1u) : 0.f);
  }
  const float* px = x + mol * xs0 + (long long)i * xs1 + (long long)c * xs2;
  float2 v;
  v.x = __ldg(px);
  v.y = second ? __ldg(px + xs2) : 0.f;
  return v;
}

// one adjacency row adj[mol, i, :] of a dense fp32 tensor, or of the bit words of a compact batch (exact 0/1 entries)
struct AdjRow {
  const float* p;
  long long s;
  const uint32_t* w;
};
__device__ __forceinline__ AdjRow adj_row(const float* __restrict__ adj, long long as0, long long as1, long long as2,
                                          const uint32_t* __restrict__ adj_bits, int aw, int N, long long mol, int i) {
  AdjRow r;
  if (adj_bits) { r.p = nullptr; r.s = 0; r.w = adj_bits + ((size_t)mol * N + i) * aw; }
  else { r.p = adj + mol * as0 + (long long)i * as1; r.s = as2; r.w = nullptr; }
  return r;
}
__device__ __forceinline__ float adj_at(const AdjRow& r, int j) {
  return r.w ? (float)((__ldg(&r.w[j >> 5]) >> (j & 31)) & 1u) : __ldg(&r.p[(long long)j * r.s]);
}

// ---- 16-byte neighbour record of one adjacency row (built once per forward by adj_pack_kernel) ----------------
// byte 0: count (bits 0-3, <= kMaxNbr) | 0x20 if the molecule's adjacency is exactly 0/1 | 0x40 if the row has a
//         zero entry (j < N) | 0x80 if count > kMaxNbr (walk the bit masks instead)
// byte 1: index of the first zero entry (the "0 * T" candidate of the reference max-pool)
// byte 2: filled when staged into shared memory: index of the molecule inside the tile (its first row is m*N)
// bytes 3..15: neighbour indices, ascending; molecule-relative in global memory, tile-relative once staged
__device__ __forceinline__ int rec_cnt(const uint4& r) { return r.x & 0xf; }
__device__ __forceinline__ bool rec_has_zero(const uint4& r) { return (r.x & 0x40) != 0; }
__device__ __forceinline__ bool rec_fallback(const uint4& r) { return (r.x & 0x80) != 0; }
__device__ __forceinline__ int rec_jz(const uint4& r) { return (r.x >> 8) & 0xff; }
__device__ __forceinline__ bool rec_binary(const uint4& r) { return (r.x & 0x20) != 0; }
__device__ __forceinline__ int rec_m(const uint4& r) { return (r.x >> 16) & 0xff; }
template <int Q>
__device__ __forceinline__ int rec_idx(const uint4& r) {
  constexpr int byte = 3 + Q;
  const unsigned w = (byte >> 2) == 0 ? r.x : (byte >> 2) == 1 ? r.y : (byte >> 2) == 2 ? r.z : r.w;
  return (int)__byte_perm(w, 0u, 0x4440u | (unsigned)(byte & 3));   // one PRMT: byte -> bits 0-7, zeros above
}
__device__ __forceinline__ uint4 rec_stage(uint4 r, int m, int mb) {   // molecule-relative -> tile-relative indices
  const unsigned add = (unsigned)mb * 0x01010101u;
  r.x = (r.x & 0x0000ffffu) | ((unsigned)m << 16) | ((((r.x >> 24) + mb) & 0xff) << 24);
  r.y = __vadd4(r.y, add);
  r.z = __vadd4(r.z, add);
  r.w = __vadd4(r.w, add);
  return r;
}
// as rec_stage, and every unused neighbour slot (>= count) points at tile row `dummy`: a row the staged buffers keep at the
// neutral element (-inf for the max-pool, 0 for the gather), so the first neighbour steps need no count checks at all
__device__ __forceinline__ uint4 rec_stage_padded(uint4 r, int m, int mb, int dummy) {
  r = rec_stage(r, m, mb);
  const int cnt = (r.x & 0x80u) ? kMaxNbr : (int)(r.x & 0xfu);
  unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int q = 0; q < kMaxNbr; ++q) {
    const int byte = 3 + q;
    if (q >= cnt) w[byte >> 2] = (w[byte >> 2] & ~(0xffu << ((byte & 3) * 8))) | ((unsigned)dummy << ((byte & 3) * 8));
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ int first_zero_bit(const u64* m, int NW, int N) {
  for (int w = 0; w < NW; ++w) {
    u64 inv = ~m[w];
    int hi = N - 64 * w;
    if (hi < 64) inv &= (hi <= 0) ? 0ull : ((1ull << hi) - 1ull);
    if (inv) return 64 * w + __ffsll((long long)inv) - 1;
  }
  return -1;
}

// ---- bit-mask neighbour loops with batched shared-memory loads -------------------------------------------
// `col` points at element (molecule base row, this lane's channel) of a [rows][stride] fp32 chunk buffer. The
// mask is warp-uniform, so index extraction runs once per warp; up to 8 neighbour values are loaded back to
// back (independent LDS) before the dependent compare / add chain consumes them in ascending-j order.
__device__ __forceinline__ void masked_max_first(const u64* mk, int NW, const float* col, int stride, float& best, int& arg, int& cnt) {
  best = -INFINITY;
  arg = 0;
  cnt = 0;
  for (int w = 0; w < NW; ++w) {
    u64 bits = mk[w];
    cnt += __popcll(bits);
    while (bits) {
      int j[8];
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        j[q] = -1;
        if (bits) { j[q] = 64 * w + __ffsll((long long)bits) - 1; bits &= bits - 1; }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = (j[q] >= 0) ? col[j[q] * stride] : -INFINITY;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (v[q] > best) { best = v[q]; arg = j[q]; }
    }
  }
}

__device__ __forceinline__ float masked_sum(const u64* mk, int NW, const float* col, int stride) {
  float s = 0.f;
  for (int w = 0; w < NW; ++w) {
    u64 bits = mk[w];
    while (bits) {
      int j[8];
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        j[q] = -1;
        if (bits) { j[q] = 64 * w + __ffsll((long long)bits) - 1; bits &= bits - 1; }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = (j[q] >= 0) ? col[j[q] * stride] : 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (j[q] >= 0) s += v[q];      // ascending j, like a dense dot product that skips exact zeros
    }
  }
  return s;
}

// tanh(x) = 1 - 2/(exp(2x)+1) on the SFU (MUFU.EX2 + MUFU.RCP): absolute error <~ 3e-7 over the whole range,
// saturates correctly (+-1) for large |x|. Same function in forward and backward.
constexpr float kTanhScale = 2.8853900817779268f;   // 2 * log2(e)
__device__ __forceinline__ float fast_tanh_scaled(float xs) {   // tanh(x) given xs = x * kTanhScale
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(xs));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return fmaf(-2.f, r, 1.f);
}
__device__ __forceinline__ float fast_tanh(float x) { return fast_tanh_scaled(x * kTanhScale); }

}  // namespace ocgcn
