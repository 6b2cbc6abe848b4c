// The molecule-tile kernel of the exact-fp32 (FFMA) mode.
//
// One CTA owns a tile of MT molecules (rows = MT*N <= TR). It streams the tile's channels in chunks of 32
// (one channel per lane): a PRODUCER turns the chunk into the left GEMM operand entirely in shared memory
// (BatchNorm apply + tanh + neighbour max-pool + A.H gather, or the backward element-wise math), the chunk is
// multiplied with the matching 32 rows of the weight matrix into fp32 register accumulators (8x8 per thread),
// and an EPILOGUE finishes the tile (bias + BatchNorm partial statistics, readout, or the A^T gather).
// Nothing but the layer boundary tensors (Z_l, arg-max, S_l) ever goes to HBM.
//
// Reference semantics implemented here: openchem/layers/gcn.py:33-42 and
// openchem/modules/encoders/gcn_encoder.py:43-53 (see include/ocgcn.h).
#pragma once
#include "ocgcn_common.cuh"

namespace ocgcn {

template <int TR>
__host__ __device__ constexpr int tile_cols() { return TR == 128 ? 128 : 64; }

// dynamic shared memory needed by tile_kernel<TR, TC>
template <int TR, int TC>
__host__ __device__ inline size_t tile_smem_bytes(int NW) {
  size_t masks = (size_t)2 * TR * NW * sizeof(u64);
  size_t main_a = (size_t)2 * TR * KCP * 4 + (size_t)KC * TC * 4;  // two chunk buffers + weight chunk
  size_t main_b = (size_t)TR * (TC + 4) * 4;                      // epilogue staging tile
  size_t mainsz = main_a > main_b ? main_a : main_b;
  return masks + mainsz + (size_t)TR * 4 /*molecule flags*/ + (size_t)kWarps * 32 * 4 /*warp partials*/ + 64;
}

template <int TR, int TC, int PROD, int EPI>
__global__ void __launch_bounds__(kThreads, (TR == 128 ? 2 : 1)) tile_kernel(const TileArgs a) {
  constexpr int TYN = TR / 8;
  constexpr int TXN = TC / 8;
  static_assert(TYN * TXN == kThreads, "thread grid must cover the tile");
  constexpr int TCP = TC + 4;
  constexpr int RPW = TR / kWarps;  // rows per warp (16 or 32), processed in batches of 16 independent loads

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.N, NW = a.NW, K = a.K, NC = a.NC;
  u64* sRow = reinterpret_cast<u64*>(smem_raw);
  u64* sCol = sRow + (size_t)TR * NW;
  float* mainp = reinterpret_cast<float*>(sCol + (size_t)TR * NW);
  float* buf0 = mainp;
  float* buf1 = buf0 + TR * KCP;
  float* sW = buf1 + TR * KCP;
  float* sC = mainp;  // epilogue staging, aliases buf0/buf1/sW
  constexpr size_t main_a = (size_t)2 * TR * KCP + (size_t)KC * TC;
  constexpr size_t main_b = (size_t)TR * TCP;
  constexpr size_t mainsz = main_a > main_b ? main_a : main_b;
  unsigned* sFlag = reinterpret_cast<unsigned*>(mainp + mainsz);
  float* sWarp = reinterpret_cast<float*>(sFlag + TR);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid % TXN, ty = tid / TXN;
  const int tile = blockIdx.x;
  const int mol0 = tile * a.MT;
  const int nm = min(a.MT, a.B - mol0);
  const int rows = nm * N;
  const size_t row0 = (size_t)mol0 * N;

  // ---- stage masks and per-molecule flags ---------------------------------------------------------------
  for (int idx = tid; idx < rows * NW; idx += kThreads) {
    sRow[idx] = a.rowmask[row0 * NW + idx];
    if (EPI == E_ATG) sCol[idx] = a.colmask[row0 * NW + idx];
  }
  for (int idx = tid; idx < nm; idx += kThreads) sFlag[idx] = a.molflag[mol0 + idx];
  __syncthreads();

  float acc[8][8];

  for (int cb = 0; cb < NC || (a.skip_gemm && cb == 0); cb += TC) {
    const bool first_pass = (cb == 0);
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

    for (int k0 = 0; k0 < K; k0 += KC) {
      const int c = k0 + lane;
      const bool cvalid = c < K;
      const float* opnd = buf0;

      // weight chunk [KC][TC] (independent of the producer; sW is free after the previous trailing sync)
      if (!a.skip_gemm) {
        for (int idx = tid; idx < KC * TC; idx += kThreads) {
          int kk = idx / TC, cc = idx - kk * TC;
          int gk = k0 + kk, gc = cb + cc;
          sW[idx] = (gk < K && gc < NC) ? __ldg(&a.Bm[(size_t)gk * NC + gc]) : 0.f;
        }
      }

      // =========================== PRODUCER ===========================
      if (PROD == P_X) {
#pragma unroll
        for (int h = 0; h < RPW / 16; ++h) {
          float v[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int r = warp + (h * 16 + q) * kWarps;
            v[q] = 0.f;
            if (r < rows && cvalid) {
              int m = r / N, i = r - m * N;
              v[q] = __ldg(&a.in0[(long long)(mol0 + m) * a.xs0 + (long long)i * a.xs1 + (long long)c * a.xs2]);
            }
          }
#pragma unroll
          for (int q = 0; q < 16; ++q) buf1[(warp + (h * 16 + q) * kWarps) * KCP + lane] = v[q];
        }
        __syncthreads();
      } else if (PROD == P_MID || PROD == P_RO) {
        const float sc = cvalid ? __ldg(&a.vec0[c]) : 0.f;
        const float sh = cvalid ? __ldg(&a.vec1[c]) : 0.f;
#pragma unroll
        for (int h = 0; h < RPW / 16; ++h) {
          float v[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {   // all loads of this warp's rows in flight before the first tanh
            const int r = warp + (h * 16 + q) * kWarps;
            v[q] = (r < rows && cvalid) ? __ldg(&a.in0[(row0 + r) * K + c]) : 0.f;
          }
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int r = warp + (h * 16 + q) * kWarps;
            buf0[r * KCP + lane] = (r < rows && cvalid) ? tanhf(fmaf(v[q], sc, sh)) : 0.f;
          }
        }
        __syncthreads();
        // neighbour max-pool: P[i,c] = max_j A[i,j]*T[j,c], first index wins ties (gcn_encoder.py:44-50)
        for (int r = warp; r < TR; r += kWarps) {
          float best = 0.f;
          int arg = 0;
          if (r < rows) {
            const int m = r / N, i = r - m * N, mb = m * N;
            if (sFlag[m]) {
              const u64* mk = &sRow[(size_t)r * NW];
              int cnt;
              masked_max_first(mk, NW, &buf0[mb * KCP + lane], KCP, best, arg, cnt);
              if (cnt < N && !(best > 0.f)) {  // a zero entry A[i,jz]=0 contributes the candidate 0*T = 0
                int jz = first_zero_bit(mk, NW, N);
                if (best < 0.f || jz < arg) arg = jz;
                best = 0.f;
              }
            } else {
              const float* arow = a.adj + (long long)(mol0 + m) * a.as0 + (long long)i * a.as1;
              best = __ldg(&arow[0]) * buf0[mb * KCP + lane];
              for (int j = 1; j < N; ++j) {
                float v = __ldg(&arow[(long long)j * a.as2]) * buf0[(mb + j) * KCP + lane];
                if (v > best) { best = v; arg = j; }
              }
            }
            if (first_pass && cvalid) {
              a.arg_out[(row0 + r) * K + c] = (unsigned char)arg;
              if (PROD == P_RO) a.prod_out[(row0 + r) * K + c] = best;
            }
          }
          buf1[r * KCP + lane] = best;
        }
        __syncthreads();
      } else if (PROD == P_DZ) {
        float g = 0.f, c1 = 0.f, c2 = 0.f, mu = 0.f, is = 0.f;
        if (cvalid) { g = __ldg(&a.vec0[c]); c1 = __ldg(&a.vec1[c]); c2 = __ldg(&a.vec2[c]); mu = __ldg(&a.vec3[c]); is = __ldg(&a.vec4[c]); }
        float colsum = 0.f;
#pragma unroll
        for (int h = 0; h < RPW / 16; ++h) {
          float zv[16], dv[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int r = warp + (h * 16 + q) * kWarps;
            const bool ok = r < rows && cvalid;
            zv[q] = ok ? __ldg(&a.in1[(row0 + r) * K + c]) : 0.f;
            dv[q] = ok ? __ldg(&a.in0[(row0 + r) * K + c]) : 0.f;
          }
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int r = warp + (h * 16 + q) * kWarps;
            float dz = 0.f;
            if (r < rows && cvalid) {
              const float yhat = (zv[q] - mu) * is;
              dz = g * (dv[q] - c1 - yhat * c2);
              if (first_pass) { a.prod_out[(row0 + r) * K + c] = dz; colsum += dz; }
            }
            buf0[r * KCP + lane] = dz;
          }
        }
        if (first_pass && a.prod_partials) sWarp[warp * 32 + lane] = colsum;
        __syncthreads();
        if (first_pass && a.prod_partials && warp == 0 && cvalid) {
          float s = 0.f;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) s += sWarp[w * 32 + lane];
          a.prod_partials[(size_t)tile * K + c] = s;
        }
      } else if (PROD == P_DU) {
        float colsum = 0.f;
#pragma unroll
        for (int h = 0; h < RPW / 16; ++h) {
          float dv[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int r = warp + (h * 16 + q) * kWarps;
            dv[q] = (r < rows && cvalid) ? __ldg(&a.in0[(row0 + r) * K + c]) : 0.f;
          }
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int r = warp + (h * 16 + q) * kWarps;
            float du = 0.f;
            if (r < rows && cvalid) {
              const int m = r / N;
              const size_t ob = (size_t)(mol0 + m) * K + c;
              const float ov = __ldg(&a.in2[ob]);
              du = __ldg(&a.in1[ob]) * (1.f - ov * ov) * (1.f - dv[q] * dv[q]);
              if (first_pass) { a.prod_out[(row0 + r) * K + c] = du; colsum += du; }
            }
            buf0[r * KCP + lane] = du;
          }
        }
        if (first_pass && a.prod_partials) sWarp[warp * 32 + lane] = colsum;
        __syncthreads();
        if (first_pass && a.prod_partials && warp == 0 && cvalid) {
          float s = 0.f;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) s += sWarp[w * 32 + lane];
          a.prod_partials[(size_t)tile * K + c] = s;
        }
      }

      if (PROD == P_X || PROD == P_MID) {
        // S = A.H for this chunk (gcn.py:34), written over the dead T buffer
        for (int r = warp; r < TR; r += kWarps) {
          float s = 0.f;
          if (r < rows) {
            const int m = r / N, i = r - m * N, mb = m * N;
            if (sFlag[m]) {
              s = masked_sum(&sRow[(size_t)r * NW], NW, &buf1[mb * KCP + lane], KCP);
            } else {
              const float* arow = a.adj + (long long)(mol0 + m) * a.as0 + (long long)i * a.as1;
              for (int j = 0; j < N; ++j) s = fmaf(__ldg(&arow[(long long)j * a.as2]), buf1[(mb + j) * KCP + lane], s);
            }
            if (first_pass && cvalid && a.prod_out) a.prod_out[(row0 + r) * K + c] = s;
          }
          buf0[r * KCP + lane] = s;
        }
        __syncthreads();
      }
      if (PROD == P_RO) opnd = buf1;

      // =========================== GEMM chunk ===========================
      if (!a.skip_gemm) {
#pragma unroll 2
        for (int kk = 0; kk < KC; kk += 4) {
          float4 av[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) av[r] = *reinterpret_cast<const float4*>(&opnd[(r * TYN + ty) * KCP + kk]);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 b0 = *reinterpret_cast<const float4*>(&sW[(kk + k4) * TC + tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sW[(kk + k4) * TC + TC / 2 + tx * 4]);
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const float ar = (k4 == 0) ? av[r].x : (k4 == 1) ? av[r].y : (k4 == 2) ? av[r].z : av[r].w;
#pragma unroll
              for (int cc = 0; cc < 8; ++cc) acc[r][cc] = fmaf(ar, bv[cc], acc[r][cc]);
            }
          }
        }
      }
      __syncthreads();  // chunk buffers and sW are free again
    }
    if (a.skip_gemm) break;

    // =========================== EPILOGUE ===========================
    // thread's output coordinates: rows r*TYN+ty, local cols {tx*4..+3, TC/2+tx*4..+3}
    if (EPI == E_ZSTATS || EPI == E_STORE) {
      float* sRed = mainp;             // [TYN][TC]
      float* sMean = mainp + TYN * TC;  // [TC]
      float colv[8];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) colv[cc] = 0.f;
      const bool vec_ok = (NC % 4 == 0);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cl = h * (TC / 2) + tx * 4;
        const int gc = cb + cl;
        float bb[4] = {0.f, 0.f, 0.f, 0.f};
        if (EPI == E_ZSTATS && a.bias) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (gc + q < NC) bb[q] = __ldg(&a.bias[gc + q]);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int row = r * TYN + ty;
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[r][h * 4 + q] += bb[q];
          if (row < rows) {
            float* dst = a.out0 + (row0 + row) * NC + gc;
            if (vec_ok && gc + 3 < NC) {
              *reinterpret_cast<float4*>(dst) = make_float4(acc[r][h * 4 + 0], acc[r][h * 4 + 1], acc[r][h * 4 + 2], acc[r][h * 4 + 3]);
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (gc + q < NC) dst[q] = acc[r][h * 4 + q];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) colv[h * 4 + q] += acc[r][h * 4 + q];
          }
        }
      }
      if (EPI == E_ZSTATS) {
        // per-tile Welford partials (mean, M2) per channel: robust against |mean| >> std, merged in fp64 later
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int q = 0; q < 4; ++q) sRed[ty * TC + h * (TC / 2) + tx * 4 + q] = colv[h * 4 + q];
        __syncthreads();
        if (tid < TC) {
          float s = 0.f;
          for (int y = 0; y < TYN; ++y) s += sRed[y * TC + tid];
          sMean[tid] = s / (float)rows;
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float mu = sMean[h * (TC / 2) + tx * 4 + q];
            float m2 = 0.f;
#pragma unroll
            for (int r = 0; r < 8; ++r)
              if (r * TYN + ty < rows) { float d = acc[r][h * 4 + q] - mu; m2 = fmaf(d, d, m2); }
            colv[h * 4 + q] = m2;
          }
        __syncthreads();  // everyone has read sRed's first use (sMean is separate)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int q = 0; q < 4; ++q) sRed[ty * TC + h * (TC / 2) + tx * 4 + q] = colv[h * 4 + q];
        __syncthreads();
        if (tid < TC && cb + tid < NC) {
          float s = 0.f;
          for (int y = 0; y < TYN; ++y) s += sRed[y * TC + tid];
          a.partials[((size_t)tile * 2 + 0) * NC + cb + tid] = sMean[tid];
          a.partials[((size_t)tile * 2 + 1) * NC + cb + tid] = s;
        }
        __syncthreads();
      }
    } else {  // E_READOUT / E_ATG: stage the finished tile in shared memory
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cl = h * (TC / 2) + tx * 4;
        const int gc = cb + cl;
        float bb[4] = {0.f, 0.f, 0.f, 0.f};
        if (EPI == E_READOUT && a.bias) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (gc + q < NC) bb[q] = __ldg(&a.bias[gc + q]);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int row = r * TYN + ty;
          float v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            v[q] = acc[r][h * 4 + q];
            if (EPI == E_READOUT) v[q] = (row < rows && gc + q < NC) ? tanhf(v[q] + bb[q]) : 0.f;   // gcn_encoder.py:51
          }
          if (EPI == E_READOUT && row < rows) {
            float* dst = a.out0 + (row0 + row) * NC + gc;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (gc + q < NC) dst[q] = v[q];
          }
          *reinterpret_cast<float4*>(&sC[row * TCP + cl]) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
      __syncthreads();
      if (EPI == E_READOUT) {
        // O[b,:] = tanh(sum over ALL N rows of D[b,i,:])  (unmasked, gcn_encoder.py:52-53)
        for (int idx = tid; idx < nm * TC; idx += kThreads) {
          const int m = idx / TC, cl = idx - m * TC;
          if (cb + cl < NC) {
            float s = 0.f;
            for (int i = 0; i < N; ++i) s += sC[(m * N + i) * TCP + cl];
            a.out1[(size_t)(mol0 + m) * NC + cb + cl] = tanhf(s);
          }
        }
      } else {
        // dHprev[j,:] = sum_i A[i,j] * dS[i,:]
        for (int r = warp; r < rows; r += kWarps) {
          const int m = r / N, j = r - m * N, mb = m * N;
          for (int cl = lane; cl < TC; cl += 32) {
            if (cb + cl >= NC) break;
            float s = 0.f;
            if (sFlag[m]) {
              s = masked_sum(&sCol[(size_t)r * NW], NW, &sC[mb * TCP + cl], TCP);
            } else {
              const float* acol = a.adj + (long long)(mol0 + m) * a.as0 + (long long)j * a.as2;
              for (int i = 0; i < N; ++i) s = fmaf(__ldg(&acol[(long long)i * a.as1]), sC[(mb + i) * TCP + cl], s);
            }
            a.out0[(row0 + r) * NC + cb + cl] = s;
          }
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace ocgcn
