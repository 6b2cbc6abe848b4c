// MLP head + loss of Graph2Label, forward and backward, exact-fp32 arithmetic.
//
// Reference: OpenChemMLP.forward (openchem/modules/mlp/openchem_mlp.py:51-61): for every layer but the last
//     Dropout -> Linear -> BatchNorm1d (over the B rows) -> activation,   last layer: Dropout -> Linear -> activation;
// criteria nn.MSELoss (mean over B*T) and MultitaskLoss (openchem/criterion/multitask_loss.py:40-49).
// In eager torch this is ~20 kernels of a few microseconds each per step; here it is one kernel per layer and direction:
//
//   head_fwd_kernel (layer i)  in: X or U_{i-1} (pre-BN)  -> [BN_{i-1} + act_{i-1} applied on load, statistics folded by every CTA from
//                              the per-tile partials of the producer: no finalize launch]  -> U_i = A_{i-1} W_i^T + b_i
//                              -> hidden: store U_i + per-tile (mean, M2);  last: predictions = act(U) + per-tile loss partials
//   head_loss_kernel           folds the loss partials (fixed order, fp64)
//   head_bwd_kernel (layer i)  dU_i (last: from the loss or an external gradient; hidden: BatchNorm backward of dY_i with dgamma / dbeta
//                              folded by every CTA from per-tile partials) -> dA_{i-1} = dU_i W_i, per-CTA partial dW_i = dU_i^T A_{i-1},
//                              db_i -> dY_{i-1} = dA_{i-1} * act'_{i-1} + per-tile (dgamma, dbeta) partials, or grad_input for layer 0
//   reduce_partials_kernel     (ocgcn_aux.cuh) merges the per-CTA dW / db partials in fixed order
//
// A tile is HR = 32 batch rows; all reductions have a fixed shape and order (no floating-point atomics): results are bit-reproducible.
// Widths up to 256, up to 4 layers. Dropout must be inactive (p = 0 or eval), as in the reference's GraphCNN configurations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ocgcn {

constexpr int HR = 32;     // rows per tile
constexpr int HT = 256;    // threads per CTA
constexpr int HKC = 32;    // reduction chunk of the GEMMs
constexpr int kHeadMaxWidth = 256;
enum HeadAct : int { HACT_ID = 0, HACT_RELU = 1, HACT_SIGMOID = 2, HACT_TANH = 3 };
enum HeadLoss : int { HLOSS_NONE = 0, HLOSS_MSE = 1, HLOSS_MULTITASK = 2 };

__device__ __forceinline__ float head_act(float u, int act) {
  switch (act) {
    case HACT_RELU: return u > 0.f ? u : 0.f;
    case HACT_SIGMOID: return 1.f / (1.f + expf(-u));
    case HACT_TANH: return tanhf(u);
    default: return u;
  }
}
// derivative expressed through the activation's OUTPUT a (what autograd's relu / sigmoid / tanh backward use)
__device__ __forceinline__ float head_act_grad(float a, int act) {
  switch (act) {
    case HACT_RELU: return a > 0.f ? 1.f : 0.f;
    case HACT_SIGMOID: return a * (1.f - a);
    case HACT_TANH: return 1.f - a * a;
    default: return 1.f;
  }
}
__host__ __device__ inline int head_pitch(int n) { return (n & 1) ? n : n + 1; }   // odd row pitch: conflict-free column access

// BatchNorm statistics of one hidden layer as a consumer kernel sees them
struct HeadBN {
  const float* part;      // [ntiles][2][C] per-tile mean and M2 of the pre-BN activations (training)
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  long long* nbt;
  float* stat;            // saved [2][C]: mean, invstd (written by CTA 0 of the consuming forward kernel, read by backward)
  int act;
};

// Every CTA folds the per-tile partial pairs part[t][0..1][c] over all tiles in the same fixed order (fp64), so all CTAs hold
// bit-identical results without a finalize launch: warp w takes tiles w, w+8, ... with its lanes across the columns (coalesced
// 128-byte rows, independent loads in flight), then the eight warp sums are added in order.
//   STATS: p0 = tile mean, p1 = tile M2 -> out0 = sum n_t*p0, out1 = sum (p1 + n_t*p0^2)   (mean = out0/B, M2 = out1 - out0^2/B)
//   else : plain sums of both
constexpr size_t kHeadFoldBytes = (size_t)(8 * 2 + 2) * kHeadMaxWidth * sizeof(double);
template <bool STATS>
__device__ __forceinline__ void head_fold(const float* __restrict__ part, int ntiles, int C, int B, double* sred, double* out0, double* out1) {
  constexpr int NJ = kHeadMaxWidth / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if ((C & 3) == 0 && C <= 128) {
    // common widths: one 16-byte load per (tile, row of the pair) and lane, eight tiles (16 loads) in flight per thread
    double b0[4] = {0.0, 0.0, 0.0, 0.0}, b1[4] = {0.0, 0.0, 0.0, 0.0};
    const int c4 = 4 * lane;
    if (c4 < C) {
      for (int t0 = warp; t0 < ntiles; t0 += 64) {
        float4 p0[8], p1[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int t = t0 + 8 * u;
          p0[u] = p1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (t < ntiles) {
            p0[u] = __ldg(reinterpret_cast<const float4*>(&part[((size_t)t * 2) * C + c4]));
            p1[u] = __ldg(reinterpret_cast<const float4*>(&part[((size_t)t * 2 + 1) * C + c4]));
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int t = t0 + 8 * u;
          if (t < ntiles) {
            const double nv = (double)min(HR, B - t * HR);
            const float q0[4] = {p0[u].x, p0[u].y, p0[u].z, p0[u].w}, q1[4] = {p1[u].x, p1[u].y, p1[u].z, p1[u].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (STATS) { b0[i] += nv * (double)q0[i]; b1[i] += (double)q1[i] + nv * (double)q0[i] * (double)q0[i]; }
              else { b0[i] += (double)q0[i]; b1[i] += (double)q1[i]; }
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sred[(warp * 2) * kHeadMaxWidth + c4 + i] = b0[i];
        sred[(warp * 2 + 1) * kHeadMaxWidth + c4 + i] = b1[i];
      }
    }
  } else {
  double a0[NJ], a1[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) a0[j] = a1[j] = 0.0;
#pragma unroll 4
  for (int t = warp; t < ntiles; t += 8) {
    const double nv = (double)min(HR, B - t * HR);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = lane + 32 * j;
      if (c < C) {
        const double p0 = (double)__ldg(&part[((size_t)t * 2) * C + c]), p1 = (double)__ldg(&part[((size_t)t * 2 + 1) * C + c]);
        if (STATS) { a0[j] += nv * p0; a1[j] += p1 + nv * p0 * p0; }
        else { a0[j] += p0; a1[j] += p1; }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = lane + 32 * j;
    sred[(warp * 2) * kHeadMaxWidth + c] = a0[j];
    sred[(warp * 2 + 1) * kHeadMaxWidth + c] = a1[j];
  }
  }
  __syncthreads();
  if (tid < C) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { s0 += sred[(w * 2) * kHeadMaxWidth + tid]; s1 += sred[(w * 2 + 1) * kHeadMaxWidth + tid]; }
    out0[tid] = s0;
    out1[tid] = s1;
  }
  __syncthreads();
}

struct HeadFwdArgs {
  const float* in;        // (B, Kin): the head's input (layer 0) or U_{i-1}
  int Kin, Nout;
  int prev_bn;            // apply BatchNorm_{i-1} + act_{i-1} to `in` while loading
  HeadBN bn;
  const float* W;         // (Nout, Kin) nn.Linear layout
  const float* b;         // (Nout) or null
  float* out;             // has_bn: U_i (pre-BN), else act(U_i) = predictions
  int has_bn;
  float* part_out;        // [ntiles][2][Nout]
  int act;                // applied when !has_bn
  int B, ntiles, training;
  float eps, momentum;
  int loss;               // last layer only
  const float* target;    // (B, Nout)
  float ignore;
  float* loss_part;       // MSE: [ntiles]; multitask: [ntiles][2][Nout] (BCE sums, valid-label counts)
};

// Register-tiled fp32 GEMM core shared by the three GEMMs of the head: both operands in shared memory with the REDUCTION index
// outermost ("k-major"), so that one 16-byte load fetches a thread's four rows (broadcast inside the warp) and one its four
// columns (consecutive lanes, conflict-free): 2 LDS.128 per 16 FFMA.
//   acc[r][4*p + c] += sum_k Ak[k][4*ty + r] * Bk[k][128*p + 4*tx + c]        r, c < 4; p < NCP (column passes of 128)
// Bk must be zero-padded to 128*NCP columns, Ak to the 32 rows.
constexpr int HAP = HR + 4;                                   // pitch of a k-major 32-row operand
__host__ __device__ inline int head_cw(int n) { return n > 128 ? 256 : 128; }   // padded column extent of a B operand
template <int NCP>
__device__ __forceinline__ void head_mma(const float* __restrict__ sAk, int apitch, const float* __restrict__ sBk, int bpitch, int K, int ty, int tx,
                                         float (&acc)[4][4 * NCP]) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 av = *reinterpret_cast<const float4*>(&sAk[k * apitch + 4 * ty]);
    const float ar[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
    for (int p = 0; p < NCP; ++p) {
      const float4 bv = *reinterpret_cast<const float4*>(&sBk[k * bpitch + 128 * p + 4 * tx]);
      const float bc[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][4 * p + c] = fmaf(ar[r], bc[c], acc[r][4 * p + c]);
    }
  }
}

static inline size_t head_fwd_smem(int Kin, int Nout) {
  const size_t gemm = ((size_t)Kin * HAP + (size_t)HKC * (head_cw(Nout) + 4) + (size_t)HR * head_pitch(Nout)) * 4;
  return (size_t)2 * kHeadMaxWidth * 4 + (gemm > kHeadFoldBytes ? gemm : kHeadFoldBytes);
}

template <int NCP>
__global__ void __launch_bounds__(HT, 2) head_fwd_kernel(const HeadFwdArgs a) {
  pdl_wait();
  extern __shared__ __align__(16) float hsm[];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int Kin = a.Kin, Nout = a.Nout, NP = head_pitch(Nout);
  constexpr int CW = 128 * NCP, BP = CW + 4;
  float* sScale = hsm;
  float* sShift = sScale + kHeadMaxWidth;
  float* sAk = sShift + kHeadMaxWidth;         // [Kin][HAP]  A_{i-1} tile, k-major
  float* sBk = sAk + (size_t)Kin * HAP;        // [HKC][BP]   chunk of W^T, zero-padded to CW columns
  float* sO = sBk + (size_t)HKC * BP;          // [HR][NP]    staged outputs for the column reductions
  if (a.prev_bn) {
    double* sred = reinterpret_cast<double*>(sAk);     // the GEMM region is free until the first tile
    double* f0 = sred + 16 * kHeadMaxWidth;
    double* f1 = f0 + kHeadMaxWidth;
    if (a.training) head_fold<true>(a.bn.part, a.ntiles, Kin, a.B, sred, f0, f1);
    if (tid < Kin) {
      const int c = tid;
      float mean, invstd;
      if (a.training) {
        const double m = f0[c] / (double)a.B;
        double m2 = f1[c] - f0[c] * m;
        m2 = m2 > 0.0 ? m2 : 0.0;
        const double var = m2 / (double)a.B;
        mean = (float)m;
        invstd = (float)(1.0 / sqrt(var + (double)a.eps));
        if (blockIdx.x == 0) {
          const float unb = (float)(m2 / (double)(a.B > 1 ? a.B - 1 : 1));
          a.bn.running_mean[c] = (1.f - a.momentum) * a.bn.running_mean[c] + a.momentum * mean;
          a.bn.running_var[c] = (1.f - a.momentum) * a.bn.running_var[c] + a.momentum * unb;
          if (c == 0 && a.bn.nbt) *a.bn.nbt += 1;
        }
      } else {
        mean = a.bn.running_mean[c];
        invstd = (float)(1.0 / sqrt((double)a.bn.running_var[c] + (double)a.eps));
      }
      if (blockIdx.x == 0) { a.bn.stat[c] = mean; a.bn.stat[Kin + c] = invstd; }
      const float sc = __ldg(&a.bn.gamma[c]) * invstd;
      sScale[c] = sc;
      sShift[c] = __ldg(&a.bn.beta[c]) - mean * sc;
    }
    __syncthreads();
  }
  const bool vecA = (Kin % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.in) & 15) == 0);
  const bool vecW = (Kin % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.W) & 15) == 0);
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int row0 = tile * HR, nv = min(HR, a.B - row0);
    // A tile, k-major in shared memory. Rows are read with 16-byte loads along k (coalesced: one row = consecutive threads) and
    // transposed by the stores; widths that are not a multiple of four (or a misaligned input) take the scalar lane = row loop.
    if (vecA) {
      const int K4 = Kin >> 2;
      for (int e = tid; e < HR * K4; e += HT) {
        const int r = e / K4, k = (e - r * K4) << 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nv) {
          v = __ldg(reinterpret_cast<const float4*>(&a.in[(size_t)(row0 + r) * Kin + k]));
          if (a.prev_bn) {
            v.x = head_act(fmaf(v.x, sScale[k], sShift[k]), a.bn.act);
            v.y = head_act(fmaf(v.y, sScale[k + 1], sShift[k + 1]), a.bn.act);
            v.z = head_act(fmaf(v.z, sScale[k + 2], sShift[k + 2]), a.bn.act);
            v.w = head_act(fmaf(v.w, sScale[k + 3], sShift[k + 3]), a.bn.act);
          }
        }
        sAk[k * HAP + r] = v.x;
        sAk[(k + 1) * HAP + r] = v.y;
        sAk[(k + 2) * HAP + r] = v.z;
        sAk[(k + 3) * HAP + r] = v.w;
      }
    } else
    for (int e = tid; e < HR * Kin; e += HT) {
      const int r = e & (HR - 1), k = e >> 5;
      float v = 0.f;
      if (r < nv) {
        v = __ldg(&a.in[(size_t)(row0 + r) * Kin + k]);
        if (a.prev_bn) v = head_act(fmaf(v, sScale[k], sShift[k]), a.bn.act);
      }
      sAk[k * HAP + r] = v;
    }
    float acc[4][4 * NCP];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4 * NCP; ++c) acc[r][c] = 0.f;
    for (int k0 = 0; k0 < Kin; k0 += HKC) {
      __syncthreads();
      if (vecW) {     // W rows (nn.Linear layout (Nout, Kin)) read 16 bytes along k, transposed by the stores; zero padding beyond Nout
        for (int e = tid; e < CW * (HKC / 4); e += HT) {
          const int o = e / (HKC / 4), kk = (e - o * (HKC / 4)) << 2;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (o < Nout && k0 + kk < Kin) v = __ldg(reinterpret_cast<const float4*>(&a.W[(size_t)o * Kin + k0 + kk]));
          sBk[kk * BP + o] = v.x;
          sBk[(kk + 1) * BP + o] = v.y;
          sBk[(kk + 2) * BP + o] = v.z;
          sBk[(kk + 3) * BP + o] = v.w;
        }
      } else
      for (int e = tid; e < HKC * CW; e += HT) {     // lane = output column: conflict-free stores, zero padding beyond Nout / Kin
        const int o = e % CW, kk = e / CW;
        sBk[kk * BP + o] = (o < Nout && k0 + kk < Kin) ? __ldg(&a.W[(size_t)o * Kin + k0 + kk]) : 0.f;
      }
      __syncthreads();
      head_mma<NCP>(sAk + (size_t)k0 * HAP, HAP, sBk, BP, min(HKC, Kin - k0), ty, tx, acc);
    }
    // epilogue: bias (+ activation), store, stage for the column reductions
#pragma unroll
    for (int p = 0; p < NCP; ++p)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int o = 128 * p + 4 * tx + c;
        if (o < Nout) {
          const float bias = a.b ? __ldg(&a.b[o]) : 0.f;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int rr = ty * 4 + r;
            float u = acc[r][4 * p + c] + bias;
            if (!a.has_bn) u = head_act(u, a.act);
            if (rr < nv) a.out[(size_t)(row0 + rr) * Nout + o] = u;
            float staged = u;
            if (!a.has_bn && a.loss != HLOSS_NONE && rr < nv) {
              const float y = __ldg(&a.target[(size_t)(row0 + rr) * Nout + o]);
              if (a.loss == HLOSS_MSE) {
                staged = (u - y) * (u - y);
              } else {   // BCE against the masked target; ignored labels count as class 0 (see head_loss_kernel)
                const float t = (y == a.ignore) ? 0.f : y;
                staged = -(t * fmaxf(logf(u), -100.f) + (1.f - t) * fmaxf(logf(1.f - u), -100.f));
              }
            }
            sO[rr * NP + o] = staged;
          }
        }
      }
    __syncthreads();
    if (a.has_bn) {
      if (tid < Nout) {     // per-tile mean and M2 over the valid rows (two passes over the staged tile)
        float s = 0.f;
        for (int r = 0; r < nv; ++r) s += sO[r * NP + tid];
        const float mean = s / (float)nv;
        float m2 = 0.f;
        for (int r = 0; r < nv; ++r) { const float d = sO[r * NP + tid] - mean; m2 = fmaf(d, d, m2); }
        a.part_out[((size_t)tile * 2) * Nout + tid] = mean;
        a.part_out[((size_t)tile * 2 + 1) * Nout + tid] = m2;
      }
    } else if (a.loss == HLOSS_MSE) {
      if (tid < 32) {       // fixed order: rows inside a column, then columns strided over the lanes, then a shuffle tree
        float s = 0.f;
        for (int o = tid; o < Nout; o += 32)
          for (int r = 0; r < nv; ++r) s += sO[r * NP + o];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (tid == 0) a.loss_part[tile] = s;
      }
    } else if (a.loss == HLOSS_MULTITASK) {
      if (tid < Nout) {
        float s = 0.f, cnt = 0.f;
        for (int r = 0; r < nv; ++r) {
          s += sO[r * NP + tid];
          cnt += (__ldg(&a.target[(size_t)(row0 + r) * Nout + tid]) == a.ignore) ? 0.f : 1.f;
        }
        a.loss_part[((size_t)tile * 2) * Nout + tid] = s;
        a.loss_part[((size_t)tile * 2 + 1) * Nout + tid] = cnt;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Narrow forward kernel: a CTA computes 32 rows x HNC = 32 output columns (grid.y walks the column blocks), so a layer of width
// 128 runs on four times as many CTAs as head_fwd_kernel, each fetching a quarter of the weight matrix in ONE round. The wide
// kernel's time is a latency chain (four weight chunks with two barriers each, cold instruction fetch) on one CTA per SM at most;
// here several CTAs share an SM and hide each other's latencies, and small batches (8 row tiles at B = 256) reach 32 SMs.
// Both operands stay row-major in shared memory with a pitch of K + 4 floats: the weight block arrives by cp.async (requested
// before the grid-dependency wait: parameters do not depend on the predecessor kernel), the A tile is a coalesced copy with
// BatchNorm + activation applied on the way. Inner loop per four k: 4 broadcast LDS.128 (this thread's rows) + 1 conflict-free
// LDS.128 (its column) for 16 FFMA, accumulated in the same k order as the wide kernel: results are bit-identical to it.
// Not used for an MSE criterion over more than 32 outputs (the per-tile loss partial needs every column in one CTA).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int HNC = 32;
static inline size_t head_fwd_narrow_smem(int Kin) {
  const size_t pa = (size_t)(Kin + 4);
  const size_t atile = (size_t)HR * pa * 4;
  return (size_t)2 * kHeadMaxWidth * 4 + (atile > kHeadFoldBytes ? atile : kHeadFoldBytes) + (size_t)HNC * pa * 4 + (size_t)HR * (HNC + 1) * 4;
}
static inline bool head_fwd_narrow_ok(int Kin, int Nout, int has_bn, int loss) {
  return (Kin % 4 == 0) && (has_bn || loss != HLOSS_MSE || Nout <= HNC);
}

__global__ void __launch_bounds__(HT, 2) head_fwd_narrow_kernel(const HeadFwdArgs a) {
  extern __shared__ __align__(16) float hsm[];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int Kin = a.Kin, Nout = a.Nout, PA = Kin + 4;
  const int o0 = blockIdx.y * HNC, o = o0 + tx;
  float* sScale = hsm;
  float* sShift = sScale + kHeadMaxWidth;
  float* sA = sShift + kHeadMaxWidth;                                       // [HR][PA]; the fold scratch aliases it before the first tile
  const size_t atile = (size_t)HR * PA * 4;
  float* sW = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sA) + (atile > kHeadFoldBytes ? atile : kHeadFoldBytes));   // [HNC][PA]
  float* sO = sW + (size_t)HNC * PA;                                        // [HR][HNC + 1]
  // weight block rows o0 .. o0 + 31 (nn.Linear layout (Nout, Kin)): straight 16-byte copies, zero rows beyond Nout
  {
    const int K4 = Kin >> 2;
    const bool aligned = (reinterpret_cast<uintptr_t>(a.W) & 15) == 0;
    for (int e = tid; e < HNC * K4; e += HT) {
      const int oc = e / K4, k = (e - oc * K4) << 2;
      float* dst = &sW[oc * PA + k];
      if (o0 + oc < Nout) {
        const float* src = &a.W[(size_t)(o0 + oc) * Kin + k];
        if (aligned) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
        } else {
          dst[0] = __ldg(src); dst[1] = __ldg(src + 1); dst[2] = __ldg(src + 2); dst[3] = __ldg(src + 3);
        }
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  pdl_wait();
  const bool lead = blockIdx.x == 0 && blockIdx.y == 0;
  if (a.prev_bn) {
    double* sred = reinterpret_cast<double*>(sA);
    double* f0 = sred + 16 * kHeadMaxWidth;
    double* f1 = f0 + kHeadMaxWidth;
    if (a.training) head_fold<true>(a.bn.part, a.ntiles, Kin, a.B, sred, f0, f1);
    if (tid < Kin) {
      const int c = tid;
      float mean, invstd;
      if (a.training) {
        const double m = f0[c] / (double)a.B;
        double m2 = f1[c] - f0[c] * m;
        m2 = m2 > 0.0 ? m2 : 0.0;
        const double var = m2 / (double)a.B;
        mean = (float)m;
        invstd = (float)(1.0 / sqrt(var + (double)a.eps));
        if (lead) {
          const float unb = (float)(m2 / (double)(a.B > 1 ? a.B - 1 : 1));
          a.bn.running_mean[c] = (1.f - a.momentum) * a.bn.running_mean[c] + a.momentum * mean;
          a.bn.running_var[c] = (1.f - a.momentum) * a.bn.running_var[c] + a.momentum * unb;
          if (c == 0 && a.bn.nbt) *a.bn.nbt += 1;
        }
      } else {
        mean = a.bn.running_mean[c];
        invstd = (float)(1.0 / sqrt((double)a.bn.running_var[c] + (double)a.eps));
      }
      if (lead) { a.bn.stat[c] = mean; a.bn.stat[Kin + c] = invstd; }
      const float sc = __ldg(&a.bn.gamma[c]) * invstd;
      sScale[c] = sc;
      sShift[c] = __ldg(&a.bn.beta[c]) - mean * sc;
    }
    __syncthreads();
  }
  const bool vecA = (reinterpret_cast<uintptr_t>(a.in) & 15) == 0;
  const float bias = (a.b && o < Nout) ? __ldg(&a.b[o]) : 0.f;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int row0 = tile * HR, nv = min(HR, a.B - row0);
    {
      const int K4 = Kin >> 2;
      for (int e = tid; e < HR * K4; e += HT) {
        const int r = e / K4, k = (e - r * K4) << 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nv) {
          const float* src = &a.in[(size_t)(row0 + r) * Kin + k];
          if (vecA) v = __ldg(reinterpret_cast<const float4*>(src));
          else v = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
          if (a.prev_bn) {
            v.x = head_act(fmaf(v.x, sScale[k], sShift[k]), a.bn.act);
            v.y = head_act(fmaf(v.y, sScale[k + 1], sShift[k + 1]), a.bn.act);
            v.z = head_act(fmaf(v.z, sScale[k + 2], sShift[k + 2]), a.bn.act);
            v.w = head_act(fmaf(v.w, sScale[k + 3], sShift[k + 3]), a.bn.act);
          }
        }
        *reinterpret_cast<float4*>(&sA[r * PA + k]) = v;
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* wrow = &sW[tx * PA];
    const float* arow = &sA[(4 * ty) * PA];
#pragma unroll 4
    for (int k = 0; k < Kin; k += 4) {
      const float4 w = *reinterpret_cast<const float4*>(&wrow[k]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 av = *reinterpret_cast<const float4*>(&arow[r * PA + k]);
        acc[r] = fmaf(av.x, w.x, acc[r]);
        acc[r] = fmaf(av.y, w.y, acc[r]);
        acc[r] = fmaf(av.z, w.z, acc[r]);
        acc[r] = fmaf(av.w, w.w, acc[r]);
      }
    }
    // epilogue: bias (+ activation), store, stage for the column reductions
    if (o < Nout) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int rr = ty * 4 + r;
        float u = acc[r] + bias;
        if (!a.has_bn) u = head_act(u, a.act);
        if (rr < nv) a.out[(size_t)(row0 + rr) * Nout + o] = u;
        float staged = u;
        if (!a.has_bn && a.loss != HLOSS_NONE && rr < nv) {
          const float y = __ldg(&a.target[(size_t)(row0 + rr) * Nout + o]);
          if (a.loss == HLOSS_MSE) {
            staged = (u - y) * (u - y);
          } else {
            const float t = (y == a.ignore) ? 0.f : y;
            staged = -(t * fmaxf(logf(u), -100.f) + (1.f - t) * fmaxf(logf(1.f - u), -100.f));
          }
        }
        sO[rr * (HNC + 1) + tx] = staged;
      }
    }
    __syncthreads();
    if (a.has_bn) {
      if (tid < HNC && o0 + tid < Nout) {     // per-tile mean and M2 over the valid rows (two passes over the staged tile)
        float s = 0.f;
        for (int r = 0; r < nv; ++r) s += sO[r * (HNC + 1) + tid];
        const float mean = s / (float)nv;
        float m2 = 0.f;
        for (int r = 0; r < nv; ++r) { const float d = sO[r * (HNC + 1) + tid] - mean; m2 = fmaf(d, d, m2); }
        a.part_out[((size_t)tile * 2) * Nout + o0 + tid] = mean;
        a.part_out[((size_t)tile * 2 + 1) * Nout + o0 + tid] = m2;
      }
    } else if (a.loss == HLOSS_MSE) {
      if (tid < 32) {       // (Nout <= 32 here) same order as the wide kernel: rows inside a column, then a shuffle tree over the columns
        float s = 0.f;
        if (tid < Nout)
          for (int r = 0; r < nv; ++r) s += sO[r * (HNC + 1) + tid];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (tid == 0) a.loss_part[tile] = s;
      }
    } else if (a.loss == HLOSS_MULTITASK) {
      if (tid < HNC && o0 + tid < Nout) {
        float s = 0.f, cnt = 0.f;
        for (int r = 0; r < nv; ++r) {
          s += sO[r * (HNC + 1) + tid];
          cnt += (__ldg(&a.target[(size_t)(row0 + r) * Nout + o0 + tid]) == a.ignore) ? 0.f : 1.f;
        }
        a.loss_part[((size_t)tile * 2) * Nout + o0 + tid] = s;
        a.loss_part[((size_t)tile * 2 + 1) * Nout + o0 + tid] = cnt;
      }
    }
    __syncthreads();
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");   // (a CTA without tiles still drains its copies before exiting)
}

// one CTA. MSE: mean over B*T. Multitask (multitask_loss.py:40-49 as executed): F.binary_cross_entropy is called with its default
// 'mean' reduction (:46), so the criterion is  s * mean_j(n_j / n_j)  with s = mean BCE over ALL B*T entries (ignored labels counted as
// class 0) and n_j = valid samples of task j — i.e. s, or nan when a task has no valid sample. coef = mean_j(n_j / n_j) is kept for
// the backward.
__global__ void __launch_bounds__(HT) head_loss_kernel(const float* __restrict__ part, int ntiles, int T, int B, int loss, float* __restrict__ out,
                                                       float* __restrict__ coef) {
  pdl_wait();
  __shared__ double sv[kHeadMaxWidth];
  __shared__ double sn[kHeadMaxWidth];
  const int tid = threadIdx.x;
  if (loss == HLOSS_MSE) {
    double s = 0.0;
    for (int t = tid; t < ntiles; t += HT) s += (double)part[t];
    sv[tid] = s;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int i = 0; i < HT; ++i) tot += sv[i];
      *out = (float)(tot / ((double)B * (double)T));
    }
  } else {
    if (tid < T) {
      double s = 0.0, n = 0.0;
      for (int t = 0; t < ntiles; ++t) {
        s += (double)part[((size_t)t * 2) * T + tid];
        n += (double)part[((size_t)t * 2 + 1) * T + tid];
      }
      sv[tid] = s;
      sn[tid] = n;
    }
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0, c = 0.0;
      for (int i = 0; i < T; ++i) { tot += sv[i]; c += (double)((float)sn[i] / (float)sn[i]); }
      const float cf = (float)(c / (double)T);
      *coef = cf;
      *out = (float)(tot / ((double)B * (double)T)) * cf;
    }
  }
}

struct HeadBwdArgs {
  int B, ntiles, training;
  const float* W;          // (Nout, Kin)
  int Nout, Kin;
  int has_bn;              // hidden layer (BatchNorm backward first) or last layer
  // hidden layer
  const float* dY;         // (B, Nout) gradient w.r.t. the BatchNorm output, written by the layer above
  const float* dpart;      // [ntiles][2][Nout] per-tile partial dgamma, dbeta
  const float* U;          // saved pre-BN activations (B, Nout)
  const float* stat;       // [2][Nout] mean, invstd
  const float* gamma;
  float* dgamma;           // final results, written by CTA 0
  float* dbeta;
  // last layer
  const float* pred;       // (B, Nout) activation output
  const float* grad_pred;  // loss == NONE: external gradient
  int act, loss;
  const float* target;
  float ignore;
  const float* gloss;      // device scalar dLoss (null = 1)
  const float* ncount;     // multitask: coef = mean_j(n_j / n_j) written by head_loss_kernel (1, or nan)
  // operand A_{i-1}
  const float* prev;       // head input (layer 0) or U_{i-1}
  int prev_bn;
  const float* pstat;
  const float* pgamma;
  const float* pbeta;
  int pact;
  // outputs
  float* dprev;            // layer 0: grad_input (may be null); else dY_{i-1} (B, Kin)
  float* dpart_out;        // [ntiles][2][Kin]
  float* dWpart;           // [gridDim.x][Nout*Kin]
  float* dbpart;           // [gridDim.x][Nout]
};

static inline int head_pad32(int n) { return (n + 31) / 32 * 32; }
static inline size_t head_bwd_smem(int Kin, int Nout) {
  const size_t gemm = ((size_t)HR * (head_pad32(Nout) + 4) + (size_t)head_pad32(Nout) * HAP + (size_t)2 * HR * (head_cw(Kin) + 4)) * 4;
  return (size_t)5 * kHeadMaxWidth * 4 + (gemm > kHeadFoldBytes ? gemm : kHeadFoldBytes);
}

template <int NCP>
__global__ void __launch_bounds__(HT, 2) head_bwd_kernel(const HeadBwdArgs a) {
  pdl_wait();
  extern __shared__ __align__(16) float hsm[];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int Kin = a.Kin, Nout = a.Nout;
  const int NO32 = (Nout + 31) / 32 * 32, NOP = NO32 + 4;
  constexpr int CW = 128 * NCP, BP = CW + 4;
  float* sC0 = hsm;                              // hidden: gamma*invstd          (per output column)
  float* sC1 = sC0 + kHeadMaxWidth;              // hidden: dbeta / B
  float* sC2 = sC1 + kHeadMaxWidth;              // hidden: dgamma / B
  float* sPs = sC2 + kHeadMaxWidth;              // BatchNorm_{i-1} scale per input column (gamma * invstd)
  float* sPt = sPs + kHeadMaxWidth;              // BatchNorm_{i-1} shift per input column (beta - mean * scale)
  float* sdUr = sPt + kHeadMaxWidth;             // [HR][NOP]   dU tile, row-major (= k-major for dW, k = row), zero-padded to NO32 columns
  float* sdUk = sdUr + (size_t)HR * NOP;         // [NO32][HAP] dU tile, k-major for dA (k = output column)
  float* sAp = sdUk + (size_t)NO32 * HAP;        // [HR][BP]    A_{i-1} tile, row-major, zero-padded to CW columns
  float* sWc = sAp + (size_t)HR * BP;            // [HKC][BP]   chunk of W rows; reused as the dY_{i-1} staging tile
  const float inv_B = 1.f / (float)a.B;
  if (a.has_bn) {
    double* sred = reinterpret_cast<double*>(sdUr);    // the GEMM region is free until the first tile
    double* f0 = sred + 16 * kHeadMaxWidth;
    double* f1 = f0 + kHeadMaxWidth;
    head_fold<false>(a.dpart, a.ntiles, Nout, a.B, sred, f0, f1);
    if (tid < Nout) {
      const int c = tid;
      const double dg = f0[c], db = f1[c];
      if (blockIdx.x == 0) { a.dgamma[c] = (float)dg; a.dbeta[c] = (float)db; }
      sC0[c] = __ldg(&a.gamma[c]) * __ldg(&a.stat[Nout + c]);
      sC1[c] = a.training ? (float)db * inv_B : 0.f;
      sC2[c] = a.training ? (float)dg * inv_B : 0.f;
    }
  }
  const float gl = a.gloss ? __ldg(a.gloss) : 1.f;
  if (a.prev_bn && tid < Kin) {     // once per CTA instead of three global loads per element of every A tile
    const float sc = __ldg(&a.pgamma[tid]) * __ldg(&a.pstat[Kin + tid]);
    sPs[tid] = sc;
    sPt[tid] = __ldg(&a.pbeta[tid]) - __ldg(&a.pstat[tid]) * sc;
  }
  __syncthreads();
  float* dWp = a.dWpart + (size_t)blockIdx.x * Nout * Kin;
  float* dbp = a.dbpart + (size_t)blockIdx.x * Nout;
  int iter = 0;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++iter) {
    const int row0 = tile * HR, nv = min(HR, a.B - row0);
    // dU tile (lane = output column: coalesced global reads), stored in both layouts
    for (int e = tid; e < HR * NO32; e += HT) {
      const int r = e / NO32, o = e - r * NO32;
      float du = 0.f;
      if (r < nv && o < Nout) {
        const size_t g = (size_t)(row0 + r) * Nout + o;
        if (a.has_bn) {
          const float yhat = (__ldg(&a.U[g]) - __ldg(&a.stat[o])) * __ldg(&a.stat[Nout + o]);
          du = sC0[o] * (__ldg(&a.dY[g]) - sC1[o] - yhat * sC2[o]);
        } else {
          const float p = __ldg(&a.pred[g]);
          float dp;
          if (a.loss == HLOSS_NONE) {
            dp = __ldg(&a.grad_pred[g]);
          } else if (a.loss == HLOSS_MSE) {
            dp = gl * 2.f * (p - __ldg(&a.target[g])) / ((float)a.B * (float)Nout);
          } else {
            const float y = __ldg(&a.target[g]);
            const float t = (y == a.ignore) ? 0.f : y;
            // d/dp of coef * mean BCE over all entries; torch's BCE backward clamps p(1-p) at 1e-12
            dp = gl * __ldg(a.ncount) * (p - t) / fmaxf((1.f - p) * p, 1e-12f) / ((float)a.B * (float)Nout);
          }
          du = dp * head_act_grad(p, a.act);
        }
      }
      sdUr[r * NOP + o] = du;
      sdUk[o * HAP + r] = du;
    }
    // A_{i-1} tile (recomputed from the saved pre-BN activations for hidden layers), zero-padded to CW columns
    for (int e = tid; e < HR * CW; e += HT) {
      const int r = e / CW, k = e - r * CW;
      float v = 0.f;
      if (r < nv && k < Kin) {
        v = __ldg(&a.prev[(size_t)(row0 + r) * Kin + k]);
        if (a.prev_bn) v = head_act(fmaf(v, sPs[k], sPt[k]), a.pact);
      }
      sAp[r * BP + k] = v;
    }
    __syncthreads();
    // db partial
    if (tid < Nout) {
      float s = 0.f;
      for (int r = 0; r < nv; ++r) s += sdUr[r * NOP + tid];
      dbp[tid] = iter ? dbp[tid] + s : s;
    }
    // dW partial, 32 output rows (o) per pass: out[o][i] = sum_r dU[r][o] * A[r][i]   (k = tile row; padded rows are zero)
    for (int ob = 0; ob < NO32; ob += 32) {
      float acc[4][4 * NCP];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4 * NCP; ++c) acc[r][c] = 0.f;
      head_mma<NCP>(sdUr + ob, NOP, sAp, BP, HR, ty, tx, acc);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int o = ob + 4 * ty + r;
        if (o < Nout) {
#pragma unroll
          for (int p = 0; p < NCP; ++p)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int i = 128 * p + 4 * tx + c;
              if (i < Kin) {
                float* dst = &dWp[(size_t)o * Kin + i];
                *dst = iter ? *dst + acc[r][4 * p + c] : acc[r][4 * p + c];
              }
            }
        }
      }
    }
    if (a.dprev) {
      // dA_{i-1}[r][i] = sum_o dU[r][o] * W[o][i]
      float acc[4][4 * NCP];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4 * NCP; ++c) acc[r][c] = 0.f;
      for (int o0 = 0; o0 < Nout; o0 += HKC) {
        __syncthreads();
        const int on = min(HKC, Nout - o0);
        for (int e = tid; e < on * CW; e += HT) {
          const int oo = e / CW, i = e - oo * CW;
          sWc[oo * BP + i] = i < Kin ? __ldg(&a.W[(size_t)(o0 + oo) * Kin + i]) : 0.f;
        }
        __syncthreads();
        head_mma<NCP>(sdUk + (size_t)o0 * HAP, HAP, sWc, BP, on, ty, tx, acc);
      }
      __syncthreads();     // all reads of sWc done: reuse it as the dY_{i-1} staging tile
#pragma unroll
      for (int p = 0; p < NCP; ++p)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = 128 * p + 4 * tx + c;
          if (i < Kin) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int rr = ty * 4 + r;
              float d = acc[r][4 * p + c];
              if (a.prev_bn) d *= head_act_grad(sAp[rr * BP + i], a.pact);
              if (rr < nv) a.dprev[(size_t)(row0 + rr) * Kin + i] = d;
              if (a.prev_bn) sWc[rr * BP + i] = (rr < nv) ? d : 0.f;
            }
          }
        }
      if (a.prev_bn) {
        __syncthreads();
        if (tid < Kin) {     // per-tile partial dgamma_{i-1} = sum_r dY * yhat, dbeta_{i-1} = sum_r dY
          const float mean = __ldg(&a.pstat[tid]), invstd = __ldg(&a.pstat[Kin + tid]);
          float dg = 0.f, db = 0.f;
          for (int r = 0; r < nv; ++r) {
            const float d = sWc[r * BP + tid];
            const float yhat = (__ldg(&a.prev[(size_t)(row0 + r) * Kin + tid]) - mean) * invstd;
            dg = fmaf(d, yhat, dg);
            db += d;
          }
          a.dpart_out[((size_t)tile * 2) * Kin + tid] = dg;
          a.dpart_out[((size_t)tile * 2 + 1) * Kin + tid] = db;
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Narrow backward kernel (same idea as head_fwd_narrow_kernel): a CTA owns 32 batch rows x HNC = 32 columns of the layer's INPUT
// width (grid.y walks the column blocks). It rebuilds the full dU tile (elementwise, cheap), keeps its [Nout][32] block of W in shared
// memory for all its tiles (cp.async, requested before the grid-dependency wait) and produces its 32 columns of dA / dY_{i-1}, of
// the dW partial and of the dgamma / dbeta partials; db comes from the CTAs of column block 0. Summation orders (over the tile rows
// for dW, over the output columns for dA) are those of head_bwd_kernel: bit-identical results.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int HNP = HNC + 4;     // pitch of the 32-column blocks in shared memory (16-byte rows, conflict-free lane = column reads)
static inline size_t head_bwd_narrow_smem(int Nout) {
  const size_t NO32 = (size_t)head_pad32(Nout), du = (size_t)HR * (NO32 + 4) * 4;
  return (size_t)3 * kHeadMaxWidth * 4 + (du > kHeadFoldBytes ? du : kHeadFoldBytes) + NO32 * HNP * 4 + (size_t)HR * HNP * 4 + (size_t)HR * (HNC + 1) * 4;
}
static inline bool head_bwd_narrow_ok(int Kin, const void* W) { return (Kin % 4 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0); }

__global__ void __launch_bounds__(HT, 2) head_bwd_narrow_kernel(const HeadBwdArgs a) {
  extern __shared__ __align__(16) float hsm[];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int Kin = a.Kin, Nout = a.Nout;
  const int NO32 = (Nout + 31) / 32 * 32, NOP = NO32 + 4;
  const int kb = blockIdx.y * HNC, kc = kb + tx;           // this thread's input column
  const bool kvalid = kc < Kin;
  float* sC0 = hsm;
  float* sC1 = sC0 + kHeadMaxWidth;
  float* sC2 = sC1 + kHeadMaxWidth;
  float* sdUr = sC2 + kHeadMaxWidth;                        // [HR][NOP] dU tile, row-major, zero-padded to NO32 columns
  const size_t du_bytes = (size_t)HR * NOP * 4;
  float* sW = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sdUr) + (du_bytes > kHeadFoldBytes ? du_bytes : kHeadFoldBytes));   // [NO32][HNP]
  float* sAp = sW + (size_t)NO32 * HNP;                     // [HR][HNP]  A_{i-1} block
  float* sSt = sAp + (size_t)HR * HNP;                      // [HR][HNC + 1] dY_{i-1} staging
  // W[:, kb .. kb + 31]: eight 16-byte pieces per output row, zero beyond Kin / Nout
  for (int e = tid; e < NO32 * (HNC / 4); e += HT) {
    const int oo = e / (HNC / 4), c4 = (e - oo * (HNC / 4)) << 2;
    float* dst = &sW[oo * HNP + c4];
    if (oo < Nout && kb + c4 < Kin)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(&a.W[(size_t)oo * Kin + kb + c4]) : "memory");
    else
      *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  pdl_wait();
  const bool lead = blockIdx.x == 0 && blockIdx.y == 0;
  const float inv_B = 1.f / (float)a.B;
  if (a.has_bn) {
    double* sred = reinterpret_cast<double*>(sdUr);
    double* f0 = sred + 16 * kHeadMaxWidth;
    double* f1 = f0 + kHeadMaxWidth;
    head_fold<false>(a.dpart, a.ntiles, Nout, a.B, sred, f0, f1);
    if (tid < Nout) {
      const int c = tid;
      const double dg = f0[c], db = f1[c];
      if (lead) { a.dgamma[c] = (float)dg; a.dbeta[c] = (float)db; }
      sC0[c] = __ldg(&a.gamma[c]) * __ldg(&a.stat[Nout + c]);
      sC1[c] = a.training ? (float)db * inv_B : 0.f;
      sC2[c] = a.training ? (float)dg * inv_B : 0.f;
    }
  }
  const float gl = a.gloss ? __ldg(a.gloss) : 1.f;
  float psc = 0.f, psh = 0.f, pmean = 0.f, pinv = 0.f;      // BatchNorm_{i-1} of this thread's input column
  if (a.prev_bn && kvalid) {
    pmean = __ldg(&a.pstat[kc]); pinv = __ldg(&a.pstat[Kin + kc]);
    psc = __ldg(&a.pgamma[kc]) * pinv;
    psh = __ldg(&a.pbeta[kc]) - pmean * psc;
  }
  __syncthreads();
  float* dWp = a.dWpart + (size_t)blockIdx.x * Nout * Kin;
  float* dbp = a.dbpart + (size_t)blockIdx.x * Nout;
  int iter = 0;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++iter) {
    const int row0 = tile * HR, nv = min(HR, a.B - row0);
    for (int e = tid; e < HR * NO32; e += HT) {
      const int r = e / NO32, o = e - r * NO32;
      float du = 0.f;
      if (r < nv && o < Nout) {
        const size_t g = (size_t)(row0 + r) * Nout + o;
        if (a.has_bn) {
          const float yhat = (__ldg(&a.U[g]) - __ldg(&a.stat[o])) * __ldg(&a.stat[Nout + o]);
          du = sC0[o] * (__ldg(&a.dY[g]) - sC1[o] - yhat * sC2[o]);
        } else {
          const float p = __ldg(&a.pred[g]);
          float dp;
          if (a.loss == HLOSS_NONE) {
            dp = __ldg(&a.grad_pred[g]);
          } else if (a.loss == HLOSS_MSE) {
            dp = gl * 2.f * (p - __ldg(&a.target[g])) / ((float)a.B * (float)Nout);
          } else {
            const float y = __ldg(&a.target[g]);
            const float t = (y == a.ignore) ? 0.f : y;
            dp = gl * __ldg(a.ncount) * (p - t) / fmaxf((1.f - p) * p, 1e-12f) / ((float)a.B * (float)Nout);
          }
          du = dp * head_act_grad(p, a.act);
        }
      }
      sdUr[r * NOP + o] = du;
    }
    // this block's 32 columns of A_{i-1}: lane = column (128 contiguous bytes per row), four rows per thread
    float uprev[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = ty + 8 * q;
      uprev[q] = (r < nv && kvalid) ? __ldg(&a.prev[(size_t)(row0 + r) * Kin + kc]) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = ty + 8 * q;
      float v = uprev[q];
      if (a.prev_bn) v = (r < nv && kvalid) ? head_act(fmaf(v, psc, psh), a.pact) : 0.f;
      sAp[r * HNP + tx] = v;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (blockIdx.y == 0 && tid < Nout) {
      float s = 0.f;
      for (int r = 0; r < nv; ++r) s += sdUr[r * NOP + tid];
      dbp[tid] = iter ? dbp[tid] + s : s;
    }
    // dW partial: out[o][kc] = sum_r dU[r][o] * A[r][kc], 32 output rows per pass (thread: four o, one column)
    for (int ob = 0; ob < NO32; ob += 32) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
      for (int r = 0; r < HR; ++r) {
        const float4 du = *reinterpret_cast<const float4*>(&sdUr[r * NOP + ob + 4 * ty]);
        const float av = sAp[r * HNP + tx];
        acc[0] = fmaf(du.x, av, acc[0]); acc[1] = fmaf(du.y, av, acc[1]); acc[2] = fmaf(du.z, av, acc[2]); acc[3] = fmaf(du.w, av, acc[3]);
      }
      if (kvalid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int o = ob + 4 * ty + j;
          if (o < Nout) {
            float* dst = &dWp[(size_t)o * Kin + kc];
            *dst = iter ? *dst + acc[j] : acc[j];
          }
        }
      }
    }
    if (a.dprev) {
      // dA_{i-1}[r][kc] = sum_o dU[r][o] * W[o][kc]   (thread: rows 4*ty .. +3, one column)
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
      for (int o = 0; o < NO32; o += 4) {
        const float w0 = sW[o * HNP + tx], w1 = sW[(o + 1) * HNP + tx], w2 = sW[(o + 2) * HNP + tx], w3 = sW[(o + 3) * HNP + tx];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 du = *reinterpret_cast<const float4*>(&sdUr[(4 * ty + r) * NOP + o]);
          acc[r] = fmaf(du.x, w0, acc[r]);
          acc[r] = fmaf(du.y, w1, acc[r]);
          acc[r] = fmaf(du.z, w2, acc[r]);
          acc[r] = fmaf(du.w, w3, acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int rr = 4 * ty + r;
        float d = acc[r];
        if (a.prev_bn) d *= head_act_grad(sAp[rr * HNP + tx], a.pact);
        if (rr < nv && kvalid) a.dprev[(size_t)(row0 + rr) * Kin + kc] = d;
        sSt[rr * (HNC + 1) + tx] = (rr < nv && kvalid) ? d : 0.f;
      }
      if (a.prev_bn) {
        __syncthreads();
        if (tid < HNC && kb + tid < Kin) {     // per-tile partial dgamma_{i-1} = sum_r dY * yhat, dbeta_{i-1} = sum_r dY
          const int k = kb + tid;
          const float mean = __ldg(&a.pstat[k]), invstd = __ldg(&a.pstat[Kin + k]);
          float dg = 0.f, db = 0.f;
          for (int r = 0; r < nv; ++r) {
            const float d = sSt[r * (HNC + 1) + tid];
            const float yhat = (__ldg(&a.prev[(size_t)(row0 + r) * Kin + k]) - mean) * invstd;
            dg = fmaf(d, yhat, dg);
            db += d;
          }
          a.dpart_out[((size_t)tile * 2) * Kin + k] = dg;
          a.dpart_out[((size_t)tile * 2 + 1) * Kin + k] = db;
        }
      }
    }
    __syncthreads();
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace ocgcn
