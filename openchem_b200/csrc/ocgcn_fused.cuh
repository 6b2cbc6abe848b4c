// Single-kernel EVAL forward of the whole GraphCNNEncoder stack (tensor-core engine): every layer of a molecule tile runs inside one
// persistent CTA and the activations never leave the SM between layers.
//
// In eval mode BatchNorm uses its running statistics (gcn.py:25,40 with self.training False), so no grid-wide reduction separates
// the layers: the accumulator of layer l (TMEM) is read back once, bias + BatchNorm(running) + tanh are applied in registers and
// T_l is written straight into the shared-memory buffers the next layer's max-pool reads. Per layer and 64-channel chunk:
//     pool(T chunk) -> H (shared)  |  S = A.H gather -> bf16 hi/lo operand images  |  tcgen05.mma (4 passes) with the chunk's weight
//     image (bulk-copied into a two-slot ring, the next piece is in flight while this one is consumed)
// then the epilogue above; the last pooled H feeds the readout GEMM directly, D = tanh(H.Wd^T + bd), O = tanh(sum_i D) (unmasked sum,
// gcn_encoder.py:51-53). Global traffic per molecule: its features, its 16-byte neighbour records, the weight images (L2) and E outputs.
//
// Same arithmetic, in the same order, as the layer-by-layer kernels (ocgcn_tile.cuh) run in eval mode: results are bit-identical
// (tests/test_gpu_fused_eval.py). Rows whose adjacency row is not a <= 13-neighbour 0/1 record (hub atoms, float adjacencies) take a
// dense scan over the molecule's columns with the reference's candidate order.
//
// One CTA per SM (about 205 KB of shared memory, 128 TMEM columns), 512 threads; launched for tiles <= a few waves per SM, where the
// layered path is bound by one tile's latency and the launch count (Tox21 / logP-sized evaluation batches).
#pragma once
#include <cooperative_groups.h>

#include "ocgcn_common.cuh"
#include "ocgcn_tc.cuh"
#include "ocgcn_tile.cuh"

namespace ocgcn {

constexpr int kFusedMaxLayers = 8;
constexpr int kFusedThreads = 512;
constexpr int kFusedBS = KC + 4;                    // padded row of the T / H chunk buffers: conflict-free 16-byte row = lane stores
constexpr int kFusedRows = 129;                     // 128 tile rows + the neutral dummy row of the padded neighbour records
constexpr int kFusedStage = 128 * (128 + 4);        // floats of the readout staging tile

// layer 0 of a compact batch: x[mol, i, c], x[mol, i, c + 1] (c = 2*lane even) of this warp's rows from the bit words. All words are
// requested first and converted afterwards, so the loads of a warp's rows overlap.
template <int RPW_, int NWARP_, int BS_>
__device__ __forceinline__ void stage_x_words(const uint32_t* __restrict__ x_bits, int xw, int N, long long mol0, const uint4* sRec, int rows,
                                              int warp, int lane, int c, bool v0, bool v1, float* sH) {
  uint32_t wv[RPW_];
#pragma unroll
  for (int q = 0; q < RPW_; ++q) {
    const int r = warp + q * NWARP_;
    wv[q] = 0u;
    if (r < rows && v0) {
      const int m = rec_m(sRec[r]), i = r - m * N;
      wv[q] = __ldg(&x_bits[((size_t)(mol0 + m) * N + i) * xw + (c >> 5)]);
    }
  }
#pragma unroll
  for (int q = 0; q < RPW_; ++q) {
    const uint32_t w = wv[q] >> (c & 31);
    *reinterpret_cast<float2*>(&sH[(warp + q * NWARP_) * BS_ + 2 * lane]) = make_float2((float)(w & 1u), v1 ? (float)((w >> 1) & 1u) : 0.f);
  }
}

struct FusedEvalArgs {
  int B, N, MT, ntiles, L;
  int F[kFusedMaxLayers + 1];                       // F[0] = input width, F[l+1] = width of layer l; all <= 128
  int E;                                            // encoder_dim <= 128
  float eps;
  const float* x;
  long long xs0, xs1, xs2;
  const float* adj;                                 // dense adjacency (rows that do not fit a neighbour record)
  long long as0, as1, as2;
  const uint32_t* x_bits;                           // compact batch (compact.py): bit words in place of x / adj (then both null)
  const uint32_t* adj_bits;
  int xw, aw;                                       // words per row: ceil(F[0]/32), ceil(N/32)
  const uint4* nbr;                                 // [B*N] row records
  const void* wimg[kFusedMaxLayers + 1];            // forward weight images (wprep_kernel), [L] = readout
  const float* bias[kFusedMaxLayers + 1];           // b_l or null, [L] = bd
  const float* gamma[kFusedMaxLayers];
  const float* beta[kFusedMaxLayers];
  const float* running_mean[kFusedMaxLayers];
  const float* running_var[kFusedMaxLayers];
  float* out;                                       // [B][E]
};

__host__ __device__ inline size_t fused_eval_smem_bytes() {
  size_t main_f = (size_t)3 * kFusedRows * kFusedBS;                          // T0, T1, H
  if ((size_t)kFusedStage > main_f) main_f = kFusedStage;
  return (size_t)128 * 16 + main_f * 4 + (size_t)2 * 8 * kImgPitch + (size_t)2 * 2 * kWPiece + (size_t)2 * kFusedMaxLayers * 128 * 4 + 128;
}

__global__ void __launch_bounds__(kFusedThreads, 1) encoder_eval_fused_kernel(const FusedEvalArgs a) {
  constexpr int NT = kFusedThreads, NWARP = NT / 32, TR = 128, BS = kFusedBS, RPW = TR / NWARP;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint4* sRec = reinterpret_cast<uint4*>(smem_raw);
  float* mainp = reinterpret_cast<float*>(sRec + TR);
  float* sT0 = mainp;
  float* sT1 = sT0 + kFusedRows * BS;
  float* sH = sT1 + kFusedRows * BS;
  float* sD = mainp;                                                          // readout staging [128][132], aliases T0 / T1 / H
  constexpr size_t main_f = ((size_t)3 * kFusedRows * kFusedBS > (size_t)kFusedStage) ? (size_t)3 * kFusedRows * kFusedBS : (size_t)kFusedStage;
  unsigned char* imgHi = reinterpret_cast<unsigned char*>(mainp + main_f);
  unsigned char* imgLo = imgHi + 8 * kImgPitch;
  unsigned char* wRing = imgLo + 8 * kImgPitch;                               // two slots of (hi, lo) pieces
  float* sScale = reinterpret_cast<float*>(wRing + 2 * 2 * kWPiece);          // [L][128] BatchNorm(running) scale * 2log2(e)
  float* sShift = sScale + kFusedMaxLayers * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sShift + kFusedMaxLayers * 128);   // [0],[1]: weight ring slots landed, [2]: MMAs complete
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int hl = lane & 15, hh = lane >> 4;
  const int N = a.N, L = a.L;
  const uint32_t img_lane_off4 = (uint32_t)(hl >> 1) * kImgPitch + (uint32_t)(hl & 1) * 8;

  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    tc::mbar_init(&bars[2], 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) {
    __syncwarp();
    tc::tmem_alloc(tslot, 128);
  }
  // dummy rows: -inf never wins a max (T buffers), 0 adds nothing (H buffer). The epilogues only ever write rows 0..127.
  for (int i = tid; i < BS; i += NT) {
    sT0[TR * BS + i] = -INFINITY;
    sT1[TR * BS + i] = -INFINITY;
    sH[TR * BS + i] = 0.f;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tslot;
  pdl_wait();
  // BatchNorm in eval mode: y = (z - running_mean) / sqrt(running_var + eps) * gamma + beta = z*scale + shift (same rounding as
  // bn_finalize_kernel's eval branch); the 2*log2(e) factor of the exp2-based tanh is folded in, as in the layered kernels
  for (int i = tid; i < L * 128; i += NT) {
    const int l = i >> 7, c = i & 127;
    float sc = 0.f, sh = 0.f;
    if (c < a.F[l + 1]) {
      const float invstd = (float)(1.0 / sqrt((double)a.running_var[l][c] + (double)a.eps));
      sc = a.gamma[l][c] * invstd;
      sh = a.beta[l][c] - (float)((double)a.running_mean[l][c]) * sc;
    }
    sScale[i] = sc * kTanhScale;
    sShift[i] = sh * kTanhScale;
  }

  uint32_t w_issued = 0, w_used = 0, mma_cnt = 0;   // weight pieces requested / consumed, commits waited for (phase parities follow)
  // the sequence of weight pieces is the same for every tile: layer-major, chunk-minor
  auto piece_src = [&](uint32_t seq) -> const unsigned char* {   // seq counts pieces within one tile
    int l = 0;
    uint32_t s = seq;
    for (;; ++l) {
      const uint32_t nch = (uint32_t)((a.F[l] + KC - 1) / KC);
      if (s < nch) break;
      s -= nch;
    }
    return reinterpret_cast<const unsigned char*>(a.wimg[l]) + (size_t)s * 2 * kWPiece;
  };
  uint32_t pieces_per_tile = 0;
  for (int l = 0; l <= L; ++l) pieces_per_tile += (uint32_t)((a.F[l] + KC - 1) / KC);
  auto issue_piece = [&]() {   // thread 0 only: next piece of this CTA's sequence -> ring slot (w_issued & 1)
    const uint32_t slot = w_issued & 1u;
    unsigned char* dst = wRing + (size_t)slot * 2 * kWPiece;
    const unsigned char* src = piece_src(w_issued % pieces_per_tile);
    tc::fence_async_smem();
    tc::mbar_expect_tx(&bars[slot], 2 * kWPiece);
    tc::bulk_g2s(dst, src, kWPiece, &bars[slot]);
    tc::bulk_g2s(dst + kWPiece, src + kWPiece, kWPiece, &bars[slot]);
  };
  const uint32_t my_tiles = ((uint32_t)a.ntiles > blockIdx.x) ? ((uint32_t)a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
  const uint32_t total_pieces = my_tiles * pieces_per_tile;
  if (tid == 0) {
    if (w_issued < total_pieces) { issue_piece(); }
  }
  if (total_pieces > 0) ++w_issued;
  if (tid == 0) {
    if (w_issued < total_pieces) { issue_piece(); }
  }
  if (w_issued < total_pieces) ++w_issued;

  auto put_operand4 = [&](int r, const float4& val) {
    uint2 hi, lo;
    tc::split4_bf16(val, hi, lo);
    const uint32_t off = img_lane_off4 + (uint32_t)r * 16;
    *reinterpret_cast<uint2*>(imgHi + off) = hi;
    *reinterpret_cast<uint2*>(imgLo + off) = lo;
  };

  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int mol0 = tile * a.MT;
    const int nm = min(a.MT, a.B - mol0);
    const int rows = nm * N;
    const size_t row0 = (size_t)mol0 * N;
    __syncthreads();   // the previous tile's readout staging (aliases T0 / T1 / H) and records are consumed
    for (int r = tid; r < TR; r += NT) {
      uint4 rec = make_uint4(0u, 0u, 0u, 0u);
      if (r < rows) {
        const int m = r / N;
        rec = rec_stage_padded(__ldg(&a.nbr[row0 + r]), m, m * N, TR);
      }
      sRec[r] = rec;
    }
    if (tid < BS) {   // the staging tile of the previous tile overwrote the dummy rows
      sT0[TR * BS + tid] = -INFINITY;
      sT1[TR * BS + tid] = -INFINITY;
      sH[TR * BS + tid] = 0.f;
    }
    __syncthreads();

    for (int l = 0; l <= L; ++l) {
      const int K = a.F[l], NC = (l < L) ? a.F[l + 1] : a.E;
      const int nch = (K + KC - 1) / KC;
      for (int ch = 0; ch < nch; ++ch) {
        const int k0 = ch * KC;
        const int c4 = k0 + 4 * hl;
        const bool cv = c4 < K;   // (K % 4 == 0 for l >= 1; layer 0 masks its tail channels when loading x)
        float* const bufT = ch ? sT1 : sT0;
        if (l == 0) {
          // H = x: strided scalar loads (two adjacent channels per lane), zero beyond the valid rows / channels
          const int c = k0 + 2 * lane;
          const bool v0 = c < K, v1 = c + 1 < K;
          if (a.x_bits) {
            stage_x_words<RPW, NWARP, BS>(a.x_bits, a.xw, N, mol0, sRec, rows, warp, lane, c, v0, v1, sH);
          } else {
#pragma unroll
            for (int q = 0; q < RPW; ++q) {
              const int r = warp + q * NWARP;
              float2 v = make_float2(0.f, 0.f);
              if (r < rows && v0) {
                const int m = rec_m(sRec[r]), i = r - m * N;
                v = load_x_pair(a.x, a.xs0, a.xs1, a.xs2, nullptr, 0, N, mol0 + m, i, c, v1);
              }
              *reinterpret_cast<float2*>(&sH[r * BS + 2 * lane]) = v;
            }
          }
          __syncthreads();
        } else {
          // neighbour max-pool of T_{l-1} (gcn_encoder.py:44-50); first index wins ties, a zero entry contributes the candidate 0
#pragma unroll 2
          for (int sI = 0; sI < RPW / 2; ++sI) {
            const int r = warp + (2 * sI + hh) * NWARP;
            float bq[4] = {0.f, 0.f, 0.f, 0.f};
            if (r < rows && cv) {
              const uint4 rec = sRec[r];
              const int mb = rec_m(rec) * N;
              if (rec_binary(rec) && !rec_fallback(rec)) {
                int aq[4];
                rec_max4_padded<BS>(rec, &bufT[4 * hl], bq, aq);
                if (rec_has_zero(rec)) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) bq[e] = fmaxf(bq[e], 0.f);
                }
              } else {   // dense scan in the reference's candidate order (float adjacency, or more neighbours than a record holds)
                const AdjRow arow = adj_row(a.adj, a.as0, a.as1, a.as2, a.adj_bits, a.aw, N, mol0 + rec_m(rec), r - mb);
                const float a00 = adj_at(arow, 0);
                const float4 t0 = *reinterpret_cast<const float4*>(&bufT[mb * BS + 4 * hl]);
                bq[0] = a00 * t0.x; bq[1] = a00 * t0.y; bq[2] = a00 * t0.z; bq[3] = a00 * t0.w;
                for (int j = 1; j < N; ++j) {
                  const float aj = adj_at(arow, j);
                  const float4 t = *reinterpret_cast<const float4*>(&bufT[(mb + j) * BS + 4 * hl]);
                  bq[0] = fmaxf(bq[0], aj * t.x); bq[1] = fmaxf(bq[1], aj * t.y);
                  bq[2] = fmaxf(bq[2], aj * t.z); bq[3] = fmaxf(bq[3], aj * t.w);
                }
              }
            }
            const float4 h4 = make_float4(bq[0], bq[1], bq[2], bq[3]);
            if (l == L) put_operand4(r, h4);                                    // readout: the pooled H is the GEMM operand
            else *reinterpret_cast<float4*>(&sH[r * BS + 4 * hl]) = h4;
          }
          if (l == L) tc::fence_async_smem();
          __syncthreads();
        }
        if (l < L) {
          // S = A.H for this chunk (gcn.py:34) -> operand images
#pragma unroll 2
          for (int sI = 0; sI < RPW / 2; ++sI) {
            const int r = warp + (2 * sI + hh) * NWARP;
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows) {
              const uint4 rec = sRec[r];
              if (rec_binary(rec) && !rec_fallback(rec)) {
                s4 = rec_sum4_padded<BS>(rec, &sH[4 * hl]);
              } else {
                const int mb = rec_m(rec) * N;
                const AdjRow arow = adj_row(a.adj, a.as0, a.as1, a.as2, a.adj_bits, a.aw, N, mol0 + rec_m(rec), r - mb);
                for (int j = 0; j < N; ++j) {
                  const float aj = adj_at(arow, j);
                  const float4 hv = *reinterpret_cast<const float4*>(&sH[(mb + j) * BS + 4 * hl]);
                  s4.x = fmaf(aj, hv.x, s4.x); s4.y = fmaf(aj, hv.y, s4.y); s4.z = fmaf(aj, hv.z, s4.z); s4.w = fmaf(aj, hv.w, s4.w);
                }
              }
            }
            put_operand4(r, s4);
          }
          tc::fence_async_smem();
          __syncthreads();
        }
        // ---- GEMM chunk: acc (+)= image . W piece, 4 passes (hi.hi + lo.hi + hi.lo + lo.lo) ----
        if (tid == 0) {
          const uint32_t slot = w_used & 1u;
          tc::mbar_wait(&bars[slot], (w_used >> 1) & 1u);
          tc::fence_after_sync();
          const int kleft = K - k0;
          const int nks = (kleft >= KC) ? KC / 16 : (kleft + 15) / 16;
          const uint32_t aHi = tc::smem_u32(imgHi), aLo = tc::smem_u32(imgLo);
          const uint32_t bHi = tc::smem_u32(wRing + (size_t)slot * 2 * kWPiece), bLo = bHi + kWPiece;
          constexpr uint32_t idesc = tc::make_idesc(tc::FMT_BF16, 128, 128, false, false);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const uint32_t ab = (p & 1) ? aLo : aHi, bb = (p & 2) ? bLo : bHi;
            uint64_t da = tc::make_desc(ab, kImgPitch, 128), db = tc::make_desc(bb, 128 * 16, 128);
            for (int ks = 0; ks < nks; ++ks) {
              tc::mma_ss<false>(tmem, da, db, idesc, (ch > 0 || p > 0 || ks > 0) ? 1u : 0u);
              da += (uint64_t)((2 * kImgPitch) >> 4);
              db += (uint64_t)((2 * 128 * 16) >> 4);
            }
          }
          tc::mma_commit(&bars[2]);
        }
        ++w_used;
        // the images, this ring slot and (last chunk) the accumulator are needed next: wait for the MMAs, then refill the slot
        if (warp == 0) tc::mbar_wait(&bars[2], mma_cnt & 1u);
        ++mma_cnt;
        __syncthreads();
        if (tid == 0) {
          if (w_issued < total_pieces) { issue_piece(); }
        }
        if (w_issued < total_pieces) ++w_issued;
      }   // chunks

      // ---- epilogue: accumulator -> registers (thread = row, 32 columns per warp group) ----
      tc::fence_after_sync();
      const int q = warp & 3, hsel = warp >> 2;
      const int row = q * 32 + lane;
      const int cl = hsel * 32;
      float v[32];
      tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cl, v);
      if (l < L) {
        // Z = acc + b; T = tanh(BatchNorm_running(Z)) -> the next layer's T chunk buffers (columns >= NC stay unused)
        float* dstT = (cl < KC ? sT0 : sT1) + row * BS + (cl & (KC - 1));
        const float* bl = a.bias[l];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float t[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int gc = cl + i + e;
            const float z = v[i + e] + ((bl && gc < NC) ? __ldg(&bl[gc]) : 0.f);
            t[e] = fast_tanh_scaled(fmaf(z, sScale[l * 128 + gc], sShift[l * 128 + gc]));
          }
          *reinterpret_cast<float4*>(dstT + i) = make_float4(t[0], t[1], t[2], t[3]);
        }
      } else {
        const float* bd = a.bias[L];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float t[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int gc = cl + i + e;
            t[e] = (row < rows && gc < NC) ? fast_tanh(v[i + e] + (bd ? __ldg(&bd[gc]) : 0.f)) : 0.f;   // gcn_encoder.py:51
          }
          *reinterpret_cast<float4*>(&sD[row * 132 + cl + i]) = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      tc::fence_before_sync();
      __syncthreads();
      if (l == L) {
        // O[b,:] = tanh(sum over ALL N rows of D[b,i,:])  (unmasked, gcn_encoder.py:52-53)
        for (int idx = tid; idx < nm * 128; idx += NT) {
          const int m = idx >> 7, c = idx & 127;
          if (c < NC) {
            float s = 0.f;
            for (int i = 0; i < N; ++i) s += sD[(m * N + i) * 132 + c];
            a.out[(size_t)(mol0 + m) * NC + c] = fast_tanh(s);
          }
        }
      }
    }   // layers
  }   // tiles
  pdl_trigger();
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}


// =====================================================================================================================
// Training-mode forward of the whole stack in ONE cooperative kernel, for batches of at most one tile per SM (the Tox21 shape, C2).
//
// Training BatchNorm needs the statistics of all B*N rows between a layer's GEMM and its tanh: a grid-wide reduction. Here the
// reduction is a cooperative grid barrier INSIDE the kernel, and the pre-BatchNorm accumulator of every tile a CTA owns stays parked
// in TMEM across it (128 columns per tile; the code handles up to four tiles per CTA, the host uses one - two tiles in sequence on
// one CTA per SM lose against two co-resident CTAs of the layer-by-layer kernels, measured at C1): after the barrier each CTA folds the per-tile partials in a
// fixed order (fp64; every CTA computes the same numbers, CTA 0 publishes them and updates the running statistics), re-reads its
// accumulators, applies bias + BatchNorm + tanh in registers and feeds the next layer's max-pool from shared memory. What the
// backward pass needs is written on the way, in exactly the formats of the layer-by-layer forward (Z_l, pool arg-max, the bf16
// operand images of S_l and H_L, D, the statistics), so ocgcn_encoder_backward runs unchanged.
// =====================================================================================================================
struct FusedTrainArgs {
  FusedEvalArgs e;
  float momentum;
  float* rm[kFusedMaxLayers];
  float* rv[kFusedMaxLayers];
  long long* nbt[kFusedMaxLayers];
  float* Z[kFusedMaxLayers];               // [rows][F[l+1]] pre-BatchNorm activations
  unsigned char* arg[kFusedMaxLayers];     // [rows][F[l+1]] pool arg-max (molecule-relative)
  void* S[kFusedMaxLayers + 1];            // bf16 tile images [tile][KP[l]/8][128][8] of S_l; [L] = H_L
  int KP[kFusedMaxLayers + 1];             // F[l] padded to a multiple of 64
  float* stats[kFusedMaxLayers];           // [4][F[l+1]] mean, invstd, scale, shift
  float* D;                                // [rows][E]
  float* partials;                         // [tiles][2][F[l+1]] per-tile (mean, M2), reused by every layer
  // stage range [l_begin, l_end) of this launch (stages 0..L-1 = layers, L = readout). The cooperative launch runs all stages with a
  // grid barrier + statistics fold after every layer; a one-stage launch (any number of tiles per CTA, no barrier) is the persistent
  // per-layer forward of large batches: its input T comes from the saved Z of the previous layer and that layer's statistics.
  int l_begin, l_end;
};

static inline size_t fused_train_smem_bytes() { return fused_eval_smem_bytes() + (size_t)2 * 4 * 128 * 4; }

__global__ void __launch_bounds__(kFusedThreads, 1) encoder_train_fused_kernel(const FusedTrainArgs t) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  const FusedEvalArgs& a = t.e;
  constexpr int NT = kFusedThreads, NWARP = NT / 32, TR = 128, BS = kFusedBS, RPW = TR / NWARP;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint4* sRec = reinterpret_cast<uint4*>(smem_raw);
  float* mainp = reinterpret_cast<float*>(sRec + TR);
  float* sT0 = mainp;
  float* sT1 = sT0 + kFusedRows * BS;
  float* sH = sT1 + kFusedRows * BS;
  float* sD = mainp;                                                          // staging tile [128][132], aliases T0 / T1 / H
  constexpr size_t main_f = ((size_t)3 * kFusedRows * kFusedBS > (size_t)kFusedStage) ? (size_t)3 * kFusedRows * kFusedBS : (size_t)kFusedStage;
  unsigned char* imgHi = reinterpret_cast<unsigned char*>(mainp + main_f);
  unsigned char* imgLo = imgHi + 8 * kImgPitch;
  unsigned char* wRing = imgLo + 8 * kImgPitch;
  float* sScale = reinterpret_cast<float*>(wRing + 2 * 2 * kWPiece);          // [128] current layer's scale * 2log2(e)  (eval layout: [L][128])
  float* sShift = sScale + kFusedMaxLayers * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sShift + kFusedMaxLayers * 128);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);
  float* sWarp = reinterpret_cast<float*>(smem_raw + fused_eval_smem_bytes());   // [2][4][128] slice statistics

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int hl = lane & 15, hh = lane >> 4;
  const int N = a.N, L = a.L;
  const uint32_t img_lane_off4 = (uint32_t)(hl >> 1) * kImgPitch + (uint32_t)(hl & 1) * 8;
  const int G = (int)gridDim.x;
  const int nk = ((int)blockIdx.x < a.ntiles) ? (a.ntiles - (int)blockIdx.x + G - 1) / G : 0;   // tiles of this CTA (<= 4)
  const bool park = (t.l_end - t.l_begin) > 1;   // several stages in one launch: accumulators stay parked in TMEM across the grid barrier
  const uint32_t tcols = (!park || nk <= 1) ? 128u : nk <= 2 ? 256u : 512u;

  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    tc::mbar_init(&bars[2], 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) {
    __syncwarp();
    tc::tmem_alloc(tslot, tcols);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tslot;
  pdl_wait();

  uint32_t w_issued = 0, w_used = 0, mma_cnt = 0;
  uint32_t total_pieces = 0;
  for (int l = t.l_begin; l < t.l_end; ++l) total_pieces += (uint32_t)(nk * ((a.F[l] + KC - 1) / KC));
  auto piece_src = [&](uint32_t seq) -> const unsigned char* {   // layer-major, then this CTA's tiles, then channel chunks
    int l = t.l_begin;
    uint32_t s = seq;
    for (;; ++l) {
      const uint32_t cnt = (uint32_t)(nk * ((a.F[l] + KC - 1) / KC));
      if (s < cnt) break;
      s -= cnt;
    }
    const uint32_t nch = (uint32_t)((a.F[l] + KC - 1) / KC);
    return reinterpret_cast<const unsigned char*>(a.wimg[l]) + (size_t)(s % nch) * 2 * kWPiece;
  };
  auto issue_piece = [&]() {   // thread 0 only
    const uint32_t slot = w_issued & 1u;
    unsigned char* dst = wRing + (size_t)slot * 2 * kWPiece;
    const unsigned char* src = piece_src(w_issued);
    tc::fence_async_smem();
    tc::mbar_expect_tx(&bars[slot], 2 * kWPiece);
    tc::bulk_g2s(dst, src, kWPiece, &bars[slot]);
    tc::bulk_g2s(dst + kWPiece, src + kWPiece, kWPiece, &bars[slot]);
  };
  for (int i = 0; i < 2; ++i) {
    if (tid == 0 && w_issued < total_pieces) issue_piece();
    if (w_issued < total_pieces) ++w_issued;
  }
  auto put_operand4 = [&](int r, const float4& val) {
    uint2 hi, lo;
    tc::split4_bf16(val, hi, lo);
    const uint32_t off = img_lane_off4 + (uint32_t)r * 16;
    *reinterpret_cast<uint2*>(imgHi + off) = hi;
    *reinterpret_cast<uint2*>(imgLo + off) = lo;
  };

  if (!park && t.l_begin > 0) {   // one-stage launch: BatchNorm scale / shift of the previous layer from its saved statistics
    const int Fp = a.F[t.l_begin];
    const float* st = t.stats[t.l_begin - 1];
    for (int c = tid; c < 128; c += NT) {
      sScale[c] = c < Fp ? __ldg(&st[2 * Fp + c]) * kTanhScale : 0.f;
      sShift[c] = c < Fp ? __ldg(&st[3 * Fp + c]) * kTanhScale : 0.f;
    }
  }
  for (int l = t.l_begin; l < t.l_end; ++l) {
    const int K = a.F[l], NC = (l < L) ? a.F[l + 1] : a.E;
    const int nch = (K + KC - 1) / KC;
    for (int k = 0; k < nk; ++k) {
      const int tile = (int)blockIdx.x + k * G;
      const int mol0 = tile * a.MT;
      const int nm = min(a.MT, a.B - mol0);
      const int rows = nm * N;
      const size_t row0 = (size_t)mol0 * N;
      const uint32_t acc = tmem + (uint32_t)(park ? k * 128 : 0);
      __syncthreads();   // the staging tile of the previous epilogue (aliases T0 / T1 / H) and the records are consumed
      for (int r = tid; r < TR; r += NT) {
        uint4 rec = make_uint4(0u, 0u, 0u, 0u);
        if (r < rows) {
          const int m = r / N;
          rec = rec_stage_padded(__ldg(&a.nbr[row0 + r]), m, m * N, TR);
        }
        sRec[r] = rec;
      }
      if (tid < BS) {
        sT0[TR * BS + tid] = -INFINITY;
        sT1[TR * BS + tid] = -INFINITY;
        sH[TR * BS + tid] = 0.f;
      }
      if (l > 0 && l == t.l_begin) {
        // one-stage launch: T = tanh(BatchNorm(Z_{l-1})) from the saved pre-BatchNorm activations (coalesced 16-byte loads, half a warp per row)
        const float* zin = t.Z[l - 1] + row0 * K;
#pragma unroll
        for (int chh = 0; chh < 2; ++chh) {
          const int c4 = chh * KC + 4 * hl;
          float* const bufT = chh ? sT1 : sT0;
          if (chh * KC < K) {
            const float4 sc4 = *reinterpret_cast<const float4*>(&sScale[c4]), sh4 = *reinterpret_cast<const float4*>(&sShift[c4]);
            float4 zv[RPW / 2];
#pragma unroll
            for (int sI = 0; sI < RPW / 2; ++sI) {
              const int r = warp + (2 * sI + hh) * NWARP;
              zv[sI] = (r < rows && c4 < K) ? __ldg(reinterpret_cast<const float4*>(&zin[r * K + c4])) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int sI = 0; sI < RPW / 2; ++sI) {
              const int r = warp + (2 * sI + hh) * NWARP;
              float4 tv = make_float4(0.f, 0.f, 0.f, 0.f);
              if (r < rows && c4 < K) {
                tv.x = fast_tanh_scaled(fmaf(zv[sI].x, sc4.x, sh4.x)); tv.y = fast_tanh_scaled(fmaf(zv[sI].y, sc4.y, sh4.y));
                tv.z = fast_tanh_scaled(fmaf(zv[sI].z, sc4.z, sh4.z)); tv.w = fast_tanh_scaled(fmaf(zv[sI].w, sc4.w, sh4.w));
              }
              *reinterpret_cast<float4*>(&bufT[r * BS + 4 * hl]) = tv;
            }
          }
        }
      } else if (l > 0) {
        // second half of layer l-1's epilogue, after the grid barrier: parked accumulator -> + bias -> BatchNorm (batch statistics)
        // -> tanh -> the T chunk buffers this layer's max-pool reads
        tc::fence_after_sync();
        const int q = warp & 3, hsel = warp >> 2;
        const int row = q * 32 + lane, cl = hsel * 32;
        float v[32];
        tc::tmem_ld32(acc + ((uint32_t)(q * 32) << 16) + (uint32_t)cl, v);
        float* dstT = (cl < KC ? sT0 : sT1) + row * BS + (cl & (KC - 1));
        const float* bl = a.bias[l - 1];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float tt[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int gc = cl + i + e;
            const float z = v[i + e] + ((bl && gc < K) ? __ldg(&bl[gc]) : 0.f);
            tt[e] = fast_tanh_scaled(fmaf(z, sScale[gc], sShift[gc]));
          }
          *reinterpret_cast<float4*>(dstT + i) = make_float4(tt[0], tt[1], tt[2], tt[3]);
        }
        tc::fence_before_sync();
      }
      __syncthreads();

      for (int ch = 0; ch < nch; ++ch) {
        const int k0 = ch * KC;
        const int c4 = k0 + 4 * hl;
        const bool cv = c4 < K;
        float* const bufT = ch ? sT1 : sT0;
        if (l == 0) {
          const int c = k0 + 2 * lane;
          const bool v0 = c < K, v1 = c + 1 < K;
          if (a.x_bits) {
            stage_x_words<RPW, NWARP, BS>(a.x_bits, a.xw, N, mol0, sRec, rows, warp, lane, c, v0, v1, sH);
          } else {
#pragma unroll
            for (int q = 0; q < RPW; ++q) {
              const int r = warp + q * NWARP;
              float2 v = make_float2(0.f, 0.f);
              if (r < rows && v0) {
                const int m = rec_m(sRec[r]), i = r - m * N;
                v = load_x_pair(a.x, a.xs0, a.xs1, a.xs2, nullptr, 0, N, mol0 + m, i, c, v1);
              }
              *reinterpret_cast<float2*>(&sH[r * BS + 2 * lane]) = v;
            }
          }
          __syncthreads();
        } else {
          // neighbour max-pool of T_{l-1} with its arg-max (the routing of the backward pass), first index wins ties
          unsigned char* argo = t.arg[l - 1] + row0 * K;
#pragma unroll 2
          for (int sI = 0; sI < RPW / 2; ++sI) {
            const int r = warp + (2 * sI + hh) * NWARP;
            float bq[4] = {0.f, 0.f, 0.f, 0.f};
            if (r < rows && cv) {
              const uint4 rec = sRec[r];
              const int mb = rec_m(rec) * N;
              int aq[4];
              if (rec_binary(rec) && !rec_fallback(rec)) {
                rec_max4_padded<BS>(rec, &bufT[4 * hl], bq, aq);
                if (rec_has_zero(rec)) {
                  const int jz = mb + rec_jz(rec);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const bool keep = bq[e] > 0.f || (bq[e] == 0.f && aq[e] < jz);
                    aq[e] = keep ? aq[e] : jz;
                    bq[e] = fmaxf(bq[e], 0.f);
                  }
                }
              } else {   // dense scan in the reference's candidate order
                const AdjRow arow = adj_row(a.adj, a.as0, a.as1, a.as2, a.adj_bits, a.aw, N, mol0 + rec_m(rec), r - mb);
                const float a00 = adj_at(arow, 0);
                const float4 t0 = *reinterpret_cast<const float4*>(&bufT[mb * BS + 4 * hl]);
                bq[0] = a00 * t0.x; bq[1] = a00 * t0.y; bq[2] = a00 * t0.z; bq[3] = a00 * t0.w;
                aq[0] = aq[1] = aq[2] = aq[3] = mb;
                for (int j = 1; j < N; ++j) {
                  const float aj = adj_at(arow, j);
                  const float4 tv = *reinterpret_cast<const float4*>(&bufT[(mb + j) * BS + 4 * hl]);
                  const float cnd[4] = {aj * tv.x, aj * tv.y, aj * tv.z, aj * tv.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    if (cnd[e] > bq[e]) { bq[e] = cnd[e]; aq[e] = mb + j; }
                }
              }
              *reinterpret_cast<uchar4*>(argo + (r * K + c4)) =
                  make_uchar4((unsigned char)(aq[0] - mb), (unsigned char)(aq[1] - mb), (unsigned char)(aq[2] - mb), (unsigned char)(aq[3] - mb));
            }
            const float4 h4 = make_float4(bq[0], bq[1], bq[2], bq[3]);
            if (l == L) put_operand4(r, h4);
            else *reinterpret_cast<float4*>(&sH[r * BS + 4 * hl]) = h4;
          }
          if (l == L) tc::fence_async_smem();
          __syncthreads();
        }
        if (l < L) {
#pragma unroll 2
          for (int sI = 0; sI < RPW / 2; ++sI) {
            const int r = warp + (2 * sI + hh) * NWARP;
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows) {
              const uint4 rec = sRec[r];
              if (rec_binary(rec) && !rec_fallback(rec)) {
                s4 = rec_sum4_padded<BS>(rec, &sH[4 * hl]);
              } else {
                const int mb = rec_m(rec) * N;
                const AdjRow arow = adj_row(a.adj, a.as0, a.as1, a.as2, a.adj_bits, a.aw, N, mol0 + rec_m(rec), r - mb);
                for (int j = 0; j < N; ++j) {
                  const float aj = adj_at(arow, j);
                  const float4 hv = *reinterpret_cast<const float4*>(&sH[(mb + j) * BS + 4 * hl]);
                  s4.x = fmaf(aj, hv.x, s4.x); s4.y = fmaf(aj, hv.y, s4.y); s4.z = fmaf(aj, hv.z, s4.z); s4.w = fmaf(aj, hv.w, s4.w);
                }
              }
            }
            put_operand4(r, s4);
          }
          tc::fence_async_smem();
          __syncthreads();
        }
        if (tid == 0) {
          // the hi image of this chunk is the formatted operand of the backward's weight-gradient product (S_l, or H_L for the readout)
          unsigned char* dst = reinterpret_cast<unsigned char*>(t.S[l]) + ((size_t)tile * (t.KP[l] / 8) + (k0 / 8)) * 2048;
          const int n8 = min(8, t.KP[l] / 8 - k0 / 8);
          for (int q = 0; q < n8; ++q) tc::bulk_s2g(dst + (size_t)q * 2048, imgHi + q * kImgPitch, 2048);
          tc::bulk_commit();
          const uint32_t slot = w_used & 1u;
          tc::mbar_wait(&bars[slot], (w_used >> 1) & 1u);
          tc::fence_after_sync();
          const int kleft = K - k0;
          const int nks = (kleft >= KC) ? KC / 16 : (kleft + 15) / 16;
          const uint32_t aHi = tc::smem_u32(imgHi), aLo = tc::smem_u32(imgLo);
          const uint32_t bHi = tc::smem_u32(wRing + (size_t)slot * 2 * kWPiece), bLo = bHi + kWPiece;
          constexpr uint32_t idesc = tc::make_idesc(tc::FMT_BF16, 128, 128, false, false);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const uint32_t ab = (p & 1) ? aLo : aHi, bb = (p & 2) ? bLo : bHi;
            uint64_t da = tc::make_desc(ab, kImgPitch, 128), db = tc::make_desc(bb, 128 * 16, 128);
            for (int ks = 0; ks < nks; ++ks) {
              tc::mma_ss<false>(acc, da, db, idesc, (ch > 0 || p > 0 || ks > 0) ? 1u : 0u);
              da += (uint64_t)((2 * kImgPitch) >> 4);
              db += (uint64_t)((2 * 128 * 16) >> 4);
            }
          }
          tc::mma_commit(&bars[2]);
        }
        ++w_used;
        if (warp == 0) tc::mbar_wait(&bars[2], mma_cnt & 1u);
        ++mma_cnt;
        if (tid == 0) tc::bulk_wait_read0();   // the image has been read by the bulk stores
        __syncthreads();
        if (tid == 0 && w_issued < total_pieces) issue_piece();
        if (w_issued < total_pieces) ++w_issued;
      }   // chunks

      // ---- first half of the epilogue: accumulator -> staging tile -> saved tensors and per-tile statistics ----
      tc::fence_after_sync();
      {
        const int q = warp & 3, hsel = warp >> 2;
        const int row = q * 32 + lane, cl = hsel * 32;
        float v[32];
        tc::tmem_ld32(acc + ((uint32_t)(q * 32) << 16) + (uint32_t)cl, v);
        if (l == L) {
          const float* bd = a.bias[L];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int gc = cl + i;
            v[i] = (row < rows && gc < NC) ? fast_tanh(v[i] + (bd ? __ldg(&bd[gc]) : 0.f)) : 0.f;   // gcn_encoder.py:51
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(&sD[row * 132 + cl + i]) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      tc::fence_before_sync();
      __syncthreads();
      {
        // coalesced store of the finished tile: Z_l = acc + b (saved for the backward), or D for the readout
        float* outp = ((l < L) ? t.Z[l] : t.D) + row0 * NC;
        const bool vec_ok = (NC % 4 == 0);
        float bb[4] = {0.f, 0.f, 0.f, 0.f};
        const int gc0 = (tid & 31) * 4;
        if (l < L && a.bias[l]) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (gc0 + i < NC) bb[i] = __ldg(&a.bias[l][gc0 + i]);
        }
        for (int idx = tid; idx < rows * 32; idx += NT) {
          const int r = idx >> 5, gc = (idx & 31) * 4;
          if (gc >= NC) continue;
          float4 v = *reinterpret_cast<const float4*>(&sD[r * 132 + gc]);
          v.x += bb[0]; v.y += bb[1]; v.z += bb[2]; v.w += bb[3];
          float* dst = outp + (r * NC + gc);
          if (vec_ok && gc + 3 < NC) {
            *reinterpret_cast<float4*>(dst) = v;
          } else {
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (gc + i < NC) dst[i] = vv[i];
          }
        }
      }
      if (l < L) {
        // per-tile (mean, M2) per channel from the staged, bias-free accumulators: four row slices, pivoted sums, fixed-order Chan merge
        const int cl = tid & 127, part = tid >> 7;
        const int r_lo = part * 32, r_hi = min(rows, r_lo + 32);
        float s = 0.f, q2 = 0.f, pv = 0.f;
        if (r_lo < r_hi) {
          pv = sD[r_lo * 132 + cl];
          for (int r = r_lo + 1; r < r_hi; ++r) { const float d = sD[r * 132 + cl] - pv; s += d; q2 = fmaf(d, d, q2); }
        }
        const float np = (float)max(r_hi - r_lo, 0);
        sWarp[part * 128 + cl] = (np > 0.f) ? pv + s / np : 0.f;
        sWarp[(4 + part) * 128 + cl] = (np > 0.f) ? fmaxf(q2 - s * s / np, 0.f) : 0.f;
        __syncthreads();
        if (tid < 128 && tid < NC) {
          float cn = 0.f, cm = 0.f, cm2 = 0.f;
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
            const float nt = (float)max(min(rows, (pp + 1) * 32) - pp * 32, 0);
            if (nt > 0.f) {
              const float totn = cn + nt, dlt = sWarp[pp * 128 + tid] - cm;
              cm += dlt * (nt / totn);
              cm2 += sWarp[(4 + pp) * 128 + tid] + dlt * dlt * (cn * nt / totn);
              cn = totn;
            }
          }
          const float bias_c = a.bias[l] ? __ldg(&a.bias[l][tid]) : 0.f;
          t.partials[((size_t)tile * 2 + 0) * NC + tid] = cm + bias_c;
          t.partials[((size_t)tile * 2 + 1) * NC + tid] = cm2;
        }
      } else {
        // O[b,:] = tanh(sum over ALL N rows of D[b,i,:])  (unmasked, gcn_encoder.py:52-53)
        for (int idx = tid; idx < nm * 128; idx += NT) {
          const int m = idx >> 7, c = idx & 127;
          if (c < NC) {
            float s = 0.f;
            for (int i = 0; i < N; ++i) s += sD[(m * N + i) * 132 + c];
            a.out[(size_t)(mol0 + m) * NC + c] = fast_tanh(s);
          }
        }
      }
    }   // tiles of this CTA

    if (l < L && l + 1 < t.l_end) {
      // ---- grid-wide BatchNorm statistics of layer l: barrier, then EVERY CTA folds all per-tile partials in the same fixed order ----
      __threadfence();
      grid.sync();
      double* sFold = reinterpret_cast<double*>(mainp);   // [4][2][128]
      {
        const int c = tid & 127, g = tid >> 7;
        double f0 = 0.0, f1 = 0.0;
        if (c < NC) {
          // eight tiles' partials in flight per round (the fold is a chain of L2 latencies otherwise); summation order stays tile-ascending
          for (int t0 = g; t0 < a.ntiles; t0 += 32) {
            float pm[8], pq[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int tt = t0 + 4 * u;
              pm[u] = tt < a.ntiles ? __ldcg(&t.partials[((size_t)tt * 2 + 0) * NC + c]) : 0.f;
              pq[u] = tt < a.ntiles ? __ldcg(&t.partials[((size_t)tt * 2 + 1) * NC + c]) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int tt = t0 + 4 * u;
              if (tt < a.ntiles) {
                const double m = (double)pm[u], nt = (double)(min(a.MT, a.B - tt * a.MT) * N);
                f0 += nt * m;
                f1 += (double)pq[u] + nt * m * m;
              }
            }
          }
        }
        sFold[(g * 2 + 0) * 128 + c] = f0;
        sFold[(g * 2 + 1) * 128 + c] = f1;
      }
      __syncthreads();
      if (tid < 128) {
        const int c = tid;
        float sc = 0.f, sh = 0.f;
        if (c < NC) {
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int g = 0; g < 4; ++g) { s0 += sFold[(g * 2 + 0) * 128 + c]; s1 += sFold[(g * 2 + 1) * 128 + c]; }
          const double n = (double)a.B * (double)N;
          const double mean = s0 / n;
          double m2 = s1 - n * mean * mean;
          m2 = m2 > 0.0 ? m2 : 0.0;
          const double var = m2 / n;
          const float invstd = (float)(1.0 / sqrt(var + (double)a.eps));
          sc = a.gamma[l][c] * invstd;
          sh = a.beta[l][c] - (float)mean * sc;
          if (blockIdx.x == 0) {
            const double unbiased = m2 / (n > 1.0 ? n - 1.0 : 1.0);
            t.rm[l][c] = (float)((1.0 - (double)t.momentum) * (double)t.rm[l][c] + (double)t.momentum * mean);
            t.rv[l][c] = (float)((1.0 - (double)t.momentum) * (double)t.rv[l][c] + (double)t.momentum * unbiased);
            if (c == 0 && t.nbt[l]) *t.nbt[l] += 1;
            float* st = t.stats[l];
            st[c] = (float)mean; st[NC + c] = invstd; st[2 * NC + c] = sc; st[3 * NC + c] = sh;
          }
        }
        sScale[c] = sc * kTanhScale;
        sShift[c] = sh * kTanhScale;
      }
      // (the next stage starts with a barrier before it touches the fold buffer's region)
    }
  }   // layers
  if (tid == 0) tc::bulk_wait0();
  pdl_trigger();
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, tcols);
}

}  // namespace ocgcn
