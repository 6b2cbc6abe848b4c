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

// every CTA folds the per-tile partials of column c in the same fixed order (fp64): bit-identical scale / shift everywhere
__device__ __forceinline__ void head_fold_stats(const float* __restrict__ part, int ntiles, int C, int B, int c, double* mean_out,
                                                double* m2_out) {
  double s = 0.0;
  for (int t = 0; t < ntiles; ++t) {
    const int nv = min(HR, B - t * HR);
    s += (double)nv * (double)__ldg(&part[((size_t)t * 2) * C + c]);
  }
  const double mean = s / (double)B;
  double m2 = 0.0;
  for (int t = 0; t < ntiles; ++t) {
    const int nv = min(HR, B - t * HR);
    const double d = (double)__ldg(&part[((size_t)t * 2) * C + c]) - mean;
    m2 += (double)__ldg(&part[((size_t)t * 2 + 1) * C + c]) + (double)nv * d * d;
  }
  *mean_out = mean;
  *m2_out = m2;
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

static inline size_t head_fwd_smem(int Kin, int Nout) {
  return ((size_t)2 * kHeadMaxWidth + (size_t)HR * head_pitch(Kin) + (size_t)HKC * head_pitch(Nout) + (size_t)HR * head_pitch(Nout)) * 4;
}

__global__ void __launch_bounds__(HT) head_fwd_kernel(const HeadFwdArgs a) {
  extern __shared__ __align__(16) float hsm[];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int Kin = a.Kin, Nout = a.Nout, KP = head_pitch(Kin), NP = head_pitch(Nout);
  float* sScale = hsm;
  float* sShift = sScale + kHeadMaxWidth;
  float* sA = sShift + kHeadMaxWidth;          // [HR][KP]
  float* sW = sA + (size_t)HR * KP;            // [HKC][NP]
  float* sO = sW + (size_t)HKC * NP;           // [HR][NP]
  if (a.prev_bn) {
    if (tid < Kin) {
      const int c = tid;
      float mean, invstd;
      if (a.training) {
        double m, m2;
        head_fold_stats(a.bn.part, a.ntiles, Kin, a.B, c, &m, &m2);
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
  constexpr int NJ = kHeadMaxWidth / 32;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int row0 = tile * HR, nv = min(HR, a.B - row0);
    for (int e = tid; e < HR * Kin; e += HT) {
      const int r = e / Kin, k = e - r * Kin;
      float v = 0.f;
      if (r < nv) {
        v = __ldg(&a.in[(size_t)(row0 + r) * Kin + k]);
        if (a.prev_bn) v = head_act(fmaf(v, sScale[k], sShift[k]), a.bn.act);
      }
      sA[r * KP + k] = v;
    }
    float acc[4][NJ];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[r][j] = 0.f;
    for (int k0 = 0; k0 < Kin; k0 += HKC) {
      __syncthreads();
      for (int e = tid; e < Nout * HKC; e += HT) {
        const int o = e / HKC, k = e - o * HKC;
        sW[k * NP + o] = (k0 + k < Kin) ? __ldg(&a.W[(size_t)o * Kin + k0 + k]) : 0.f;
      }
      __syncthreads();
      const int kn = min(HKC, Kin - k0);
      for (int k = 0; k < kn; ++k) {
        float av[4], wv[NJ];
#pragma unroll
        for (int r = 0; r < 4; ++r) av[r] = sA[(ty * 4 + r) * KP + k0 + k];
#pragma unroll
        for (int j = 0; j < NJ; ++j) wv[j] = (tx + 32 * j < Nout) ? sW[k * NP + tx + 32 * j] : 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < NJ; ++j) acc[r][j] = fmaf(av[r], wv[j], acc[r][j]);
      }
    }
    // epilogue: bias (+ activation), store, stage for the column reductions
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int o = tx + 32 * j;
      if (o < Nout) {
        const float bias = a.b ? __ldg(&a.b[o]) : 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int rr = ty * 4 + r;
          float u = acc[r][j] + bias;
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

// one CTA. MSE: mean over B*T. Multitask (multitask_loss.py:40-49 as executed): F.binary_cross_entropy is called with its default
// 'mean' reduction (:46), so the criterion is  s * mean_j(n_j / n_j)  with s = mean BCE over ALL B*T entries (ignored labels counted as
// class 0) and n_j = valid samples of task j — i.e. s, or nan when a task has no valid sample. coef = mean_j(n_j / n_j) is kept for
// the backward.
__global__ void __launch_bounds__(HT) head_loss_kernel(const float* __restrict__ part, int ntiles, int T, int B, int loss, float* __restrict__ out,
                                                       float* __restrict__ coef) {
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

static inline size_t head_bwd_smem(int Kin, int Nout) {
  return ((size_t)4 * kHeadMaxWidth + (size_t)HR * head_pitch(Nout) + (size_t)HR * head_pitch(Kin) + (size_t)HKC * head_pitch(Kin)) * 4;
}

__global__ void __launch_bounds__(HT) head_bwd_kernel(const HeadBwdArgs a) {
  extern __shared__ __align__(16) float hsm[];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int Kin = a.Kin, Nout = a.Nout, KP = head_pitch(Kin), NP = head_pitch(Nout);
  float* sC0 = hsm;                              // hidden: gamma*invstd          (per output column)
  float* sC1 = sC0 + kHeadMaxWidth;              // hidden: dbeta / B
  float* sC2 = sC1 + kHeadMaxWidth;              // hidden: dgamma / B
  float* sPs = sC2 + kHeadMaxWidth;              // previous layer: scale (first Kin) ... shift kept in registers-free form below
  float* sdU = sPs + kHeadMaxWidth;              // [HR][NP]
  float* sAp = sdU + (size_t)HR * NP;            // [HR][KP]  A_{i-1}
  float* sWc = sAp + (size_t)HR * KP;            // [HKC][KP] chunk of W rows; reused as the dY_{i-1} staging tile [HR][KP]
  const float inv_B = 1.f / (float)a.B;
  if (a.has_bn && tid < Nout) {
    const int c = tid;
    double dg = 0.0, db = 0.0;
    for (int t = 0; t < a.ntiles; ++t) {
      dg += (double)__ldg(&a.dpart[((size_t)t * 2) * Nout + c]);
      db += (double)__ldg(&a.dpart[((size_t)t * 2 + 1) * Nout + c]);
    }
    if (blockIdx.x == 0) { a.dgamma[c] = (float)dg; a.dbeta[c] = (float)db; }
    sC0[c] = __ldg(&a.gamma[c]) * __ldg(&a.stat[Nout + c]);
    sC1[c] = a.training ? (float)db * inv_B : 0.f;
    sC2[c] = a.training ? (float)dg * inv_B : 0.f;
  }
  const float gl = a.gloss ? __ldg(a.gloss) : 1.f;
  __syncthreads();
  constexpr int NJ = kHeadMaxWidth / 32;
  float* dWp = a.dWpart + (size_t)blockIdx.x * Nout * Kin;
  float* dbp = a.dbpart + (size_t)blockIdx.x * Nout;
  int iter = 0;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++iter) {
    const int row0 = tile * HR, nv = min(HR, a.B - row0);
    // dU tile
    for (int e = tid; e < HR * Nout; e += HT) {
      const int r = e / Nout, o = e - r * Nout;
      float du = 0.f;
      if (r < nv) {
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
      sdU[r * NP + o] = du;
    }
    // A_{i-1} tile (recomputed from the saved pre-BN activations for hidden layers)
    for (int e = tid; e < HR * Kin; e += HT) {
      const int r = e / Kin, k = e - r * Kin;
      float v = 0.f;
      if (r < nv) {
        v = __ldg(&a.prev[(size_t)(row0 + r) * Kin + k]);
        if (a.prev_bn) {
          const float sc = __ldg(&a.pgamma[k]) * __ldg(&a.pstat[Kin + k]);
          v = head_act(fmaf(v, sc, __ldg(&a.pbeta[k]) - __ldg(&a.pstat[k]) * sc), a.pact);
        }
      }
      sAp[r * KP + k] = v;
    }
    __syncthreads();
    // db partial
    if (tid < Nout) {
      float s = 0.f;
      for (int r = 0; r < nv; ++r) s += sdU[r * NP + tid];
      dbp[tid] = iter ? dbp[tid] + s : s;
    }
    // dW partial: out[o][i] = sum_r dU[r][o] * A[r][i]; o = ty + 8*(4*pass + q), i = tx + 32*j
    for (int pass = 0; pass * 32 < Nout; ++pass) {
      float acc[4][NJ];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[q][j] = 0.f;
      for (int r = 0; r < nv; ++r) {
        float dv[4], av[NJ];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int o = ty + 8 * (4 * pass + q);
          dv[q] = o < Nout ? sdU[r * NP + o] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) av[j] = (tx + 32 * j < Kin) ? sAp[r * KP + tx + 32 * j] : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int j = 0; j < NJ; ++j) acc[q][j] = fmaf(dv[q], av[j], acc[q][j]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int o = ty + 8 * (4 * pass + q);
        if (o < Nout) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int i = tx + 32 * j;
            if (i < Kin) {
              float* dst = &dWp[(size_t)o * Kin + i];
              *dst = iter ? *dst + acc[q][j] : acc[q][j];
            }
          }
        }
      }
    }
    if (a.dprev) {
      // dA_{i-1}[r][i] = sum_o dU[r][o] * W[o][i]
      float acc[4][NJ];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[r][j] = 0.f;
      for (int o0 = 0; o0 < Nout; o0 += HKC) {
        __syncthreads();
        const int on = min(HKC, Nout - o0);
        for (int e = tid; e < on * Kin; e += HT) {
          const int oo = e / Kin, i = e - oo * Kin;
          sWc[oo * KP + i] = __ldg(&a.W[(size_t)(o0 + oo) * Kin + i]);
        }
        __syncthreads();
        for (int oo = 0; oo < on; ++oo) {
          float dv[4], wv[NJ];
#pragma unroll
          for (int r = 0; r < 4; ++r) dv[r] = sdU[(ty * 4 + r) * NP + o0 + oo];
#pragma unroll
          for (int j = 0; j < NJ; ++j) wv[j] = (tx + 32 * j < Kin) ? sWc[oo * KP + tx + 32 * j] : 0.f;
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[r][j] = fmaf(dv[r], wv[j], acc[r][j]);
        }
      }
      __syncthreads();     // all reads of sWc done: reuse it as the dY_{i-1} staging tile
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int i = tx + 32 * j;
        if (i < Kin) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int rr = ty * 4 + r;
            float d = acc[r][j];
            if (a.prev_bn) d *= head_act_grad(sAp[rr * KP + i], a.pact);
            if (rr < nv) a.dprev[(size_t)(row0 + rr) * Kin + i] = d;
            if (a.prev_bn) sWc[rr * KP + i] = (rr < nv) ? d : 0.f;
          }
        }
      }
      if (a.prev_bn) {
        __syncthreads();
        if (tid < Kin) {     // per-tile partial dgamma_{i-1} = sum_r dY * yhat, dbeta_{i-1} = sum_r dY
          const float mean = __ldg(&a.pstat[tid]), invstd = __ldg(&a.pstat[Kin + tid]);
          float dg = 0.f, db = 0.f;
          for (int r = 0; r < nv; ++r) {
            const float d = sWc[r * KP + tid];
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

}  // namespace ocgcn
