// Streaming / reduction kernels around the tile kernel: adjacency packing, BatchNorm finalisation,
// pool-scatter + tanh' (backward first half), the split-K weight-gradient GEMM and fixed-order reductions.
#pragma once
#include "ocgcn_common.cuh"

namespace ocgcn {

// ---------------------------------------------------------------------------------------------------------
// adjacency -> bit masks. One CTA per molecule. Also classifies the molecule: molflag = 1 iff every entry is
// exactly 0 or 1 (what openchem/utils/graph.py:77-86 produces); other molecules take the dense float path.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 build_record(const u64* mk, int NW, int N, bool binary) {
  u64 lo = 0, hi = 0;   // bytes 0-7, 8-15 of the record
  int cnt = 0;
  for (int w = 0; w < NW; ++w) {
    u64 bits = mk[w];
    while (bits) {
      const u64 j = (u64)(64 * w + __ffsll((long long)bits) - 1);
      bits &= bits - 1;
      if (cnt < kMaxNbr) {
        const int k = 3 + cnt;
        if (k < 8) lo |= j << (8 * k);
        else hi |= j << (8 * (k - 8));
      }
      ++cnt;
    }
  }
  const int jz = first_zero_bit(mk, NW, N);
  u64 flags = (cnt > kMaxNbr) ? 0x80ull : (u64)cnt;
  if (jz >= 0) { flags |= 0x40ull; lo |= (u64)jz << 8; }
  if (binary) flags |= 0x20ull;
  lo |= flags;
  return make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)hi, (unsigned)(hi >> 32));
}

// BITS: the adjacency arrives as the bit words of a compact batch (compact.py: bit (j & 31) of word adj_bits[(b*N + i)*AW + (j >> 5)],
// AW = ceil(N/32)); the row masks are those words, and every entry is 0 or 1 by construction. Otherwise `adj` is dense fp32.
template <bool BITS>
__global__ void __launch_bounds__(kThreads) adj_pack_kernel(const float* __restrict__ adj, long long as0, long long as1,
                                                            long long as2, const uint32_t* __restrict__ adj_bits, int N, int NW,
                                                            u64* __restrict__ rowmask, u64* __restrict__ colmask,
                                                            unsigned* __restrict__ molflag, uint4* __restrict__ nbr,
                                                            uint4* __restrict__ cnbr) {
  pdl_wait();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* sR = reinterpret_cast<u64*>(smem_raw);
  u64* sCm = sR + (size_t)N * NW;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* ab = adj + (long long)b * as0;
  int nonbin = 0;
  if (BITS) {
    const int AW = (N + 31) >> 5;
    const uint32_t* wb = adj_bits + (size_t)b * N * AW;
    const uint32_t tail = (N & 31) ? ((1u << (N & 31)) - 1u) : 0xffffffffu;    // bits beyond column N - 1 are ignored
    for (int idx = tid; idx < N * NW; idx += kThreads) {
      const int i = idx / NW, w = idx - i * NW;
      uint32_t lo = 0u, hi = 0u;
      if (2 * w < AW) { lo = __ldg(&wb[i * AW + 2 * w]); if (2 * w == AW - 1) lo &= tail; }
      if (2 * w + 1 < AW) { hi = __ldg(&wb[i * AW + 2 * w + 1]); if (2 * w + 1 == AW - 1) hi &= tail; }
      const u64 m = (u64)lo | ((u64)hi << 32);
      sR[idx] = m;
      rowmask[(size_t)b * N * NW + idx] = m;
    }
  } else
  // Row masks. Rows in batches of four per warp: the 8*NW loads of a batch are in flight together (the kernel is latency-bound
  // otherwise). The ballots of a batch are parked in lanes 0 .. 4*NW-1 (slot = u*NW + w) and written with ONE store each to
  // shared and global memory per batch instead of one predicated lane-0 store per word.
  for (int i0 = warp; i0 < N; i0 += 4 * kWarps) {
    u64 mine = 0;
    for (int w = 0; w < NW; ++w) {
      const int j0 = 64 * w + lane, j1 = j0 + 32;
      const long long o0 = (long long)j0 * as2, o1 = (long long)j1 * as2;
      float v0[4], v1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kWarps;
        const float* rp = ab + (long long)i * as1;
        v0[u] = (i < N && j0 < N) ? __ldg(rp + o0) : 0.f;
        v1[u] = (i < N && j1 < N) ? __ldg(rp + o1) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        // v*(v-1) == 0 exactly for v in {0, 1} (and -0), non-zero (or NaN) for everything else
        nonbin |= (fabsf(v0[u] * (v0[u] - 1.f)) + fabsf(v1[u] * (v1[u] - 1.f)) != 0.f);
        const unsigned b0 = __ballot_sync(0xffffffffu, v0[u] != 0.f), b1 = __ballot_sync(0xffffffffu, v1[u] != 0.f);
        if (lane == u * NW + w) mine = (u64)b0 | ((u64)b1 << 32);
      }
    }
    if (lane < 4 * NW) {
      const int u = lane / NW, w = lane - u * NW, i = i0 + u * kWarps;
      if (i < N) {
        sR[i * NW + w] = mine;
        rowmask[((size_t)b * N + i) * NW + w] = mine;
      }
    }
  }
  nonbin = __syncthreads_or(nonbin);
  // Column masks = bit transpose of the row masks, by ballots. A warp takes (row group g of 32 rows, block of 16 columns): every
  // lane fetches the 16 bits of its row ONCE, the 16 votes are parked in lanes 0..15 and written as the g-th 32-bit half-word of
  // the column's mask row with one store (little-endian: half-word g of a row of NW 64-bit words).
  {
    const int RG = (N + 31) >> 5, CB = (N + 15) >> 4;
    unsigned* sCm32 = reinterpret_cast<unsigned*>(sCm);
    unsigned* col32 = reinterpret_cast<unsigned*>(colmask) + (size_t)b * N * NW * 2;
    for (int combo = warp; combo < RG * CB; combo += kWarps) {
      const int g = combo / CB, c0 = (combo - g * CB) << 4;
      const int i = 32 * g + lane;
      const unsigned bits16 = i < N ? (unsigned)(sR[i * NW + (c0 >> 6)] >> (c0 & 63)) & 0xffffu : 0u;
      unsigned parked = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const unsigned v = __ballot_sync(0xffffffffu, (bits16 >> k) & 1u);
        if (lane == k) parked = v;
      }
      const int j = c0 + lane;
      if (lane < 16 && j < N) {
        sCm32[j * NW * 2 + g] = parked;
        col32[(size_t)j * NW * 2 + g] = parked;
      }
    }
    // rows N .. 64*NW-1 do not exist: the upper half-words of the last mask word that no row group covered are zero
    for (int idx = tid; idx < N * (NW * 2 - RG); idx += kThreads) {
      const int j = idx / (NW * 2 - RG), h = RG + idx - j * (NW * 2 - RG);
      sCm32[j * NW * 2 + h] = 0u;
      col32[(size_t)j * NW * 2 + h] = 0u;
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * N; i += kThreads) {   // one 16-byte neighbour record per row and per column
    if (i < N) nbr[(size_t)b * N + i] = build_record(&sR[i * NW], NW, N, !nonbin);
    else cnbr[(size_t)b * N + (i - N)] = build_record(&sCm[(i - N) * NW], NW, N, !nonbin);
  }
  if (tid == 0) molflag[b] = nonbin ? 0u : 1u;
}

// N <= 64 (one 64-bit mask word per row, NW == 1): one WARP per molecule, no shared memory, no block barrier. Lane t owns rows /
// columns t and t + 32: the row masks are read (bit words) or voted (dense fp32, eight rows = 16 loads in flight per lane), the column
// masks are their bit transpose by 64 ballot pairs, and every lane builds the records of its own rows and columns. Same outputs as
// adj_pack_kernel, bit for bit; the molecule-per-CTA kernel is latency-bound at these sizes (three barriers for ~1 KB of work).
template <bool BITS>
__global__ void __launch_bounds__(kThreads) adj_pack_warp_kernel(const float* __restrict__ adj, long long as0, long long as1,
                                                                 long long as2, const uint32_t* __restrict__ adj_bits, int B, int N,
                                                                 u64* __restrict__ rowmask, u64* __restrict__ colmask,
                                                                 unsigned* __restrict__ molflag, uint4* __restrict__ nbr,
                                                                 uint4* __restrict__ cnbr) {
  pdl_wait();
  const int lane = threadIdx.x & 31, b = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (b >= B) return;   // warp-uniform
  u64 rm0 = 0, rm1 = 0;
  int nonbin = 0;
  if (BITS) {
    const int AW = (N + 31) >> 5;
    const uint32_t* wb = adj_bits + (size_t)b * N * AW;
    const uint32_t tail = (N & 31) ? ((1u << (N & 31)) - 1u) : 0xffffffffu;    // bits beyond column N - 1 are ignored
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lane + 32 * h;
      u64 m = 0;
      if (r < N) {
        uint32_t lo = __ldg(&wb[r * AW]), hi = 0u;
        if (AW == 1) lo &= tail;
        else hi = __ldg(&wb[r * AW + 1]) & tail;
        m = (u64)lo | ((u64)hi << 32);
      }
      if (h == 0) rm0 = m; else rm1 = m;
    }
  } else {
    const float* ab = adj + (long long)b * as0;
    const long long o0 = (long long)lane * as2, o1 = (long long)(lane + 32) * as2;
    for (int i0 = 0; i0 < N; i0 += 8) {
      float v0[8], v1[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u;
        const float* rp = ab + (long long)i * as1;
        v0[u] = (i < N && lane < N) ? __ldg(rp + o0) : 0.f;
        v1[u] = (i < N && lane + 32 < N) ? __ldg(rp + o1) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u;
        nonbin |= (fabsf(v0[u] * (v0[u] - 1.f)) + fabsf(v1[u] * (v1[u] - 1.f)) != 0.f);
        const unsigned b0 = __ballot_sync(0xffffffffu, v0[u] != 0.f), b1 = __ballot_sync(0xffffffffu, v1[u] != 0.f);
        const u64 m = (u64)b0 | ((u64)b1 << 32);
        if (lane == (i & 31)) { if (i < 32) rm0 = m; else rm1 = m; }
      }
    }
    nonbin = __any_sync(0xffffffffu, nonbin);
  }
  u64 cm0 = 0, cm1 = 0;
#pragma unroll 8
  for (int j = 0; j < 32; ++j) {
    const unsigned c0 = __ballot_sync(0xffffffffu, (unsigned)(rm0 >> j) & 1u), c1 = __ballot_sync(0xffffffffu, (unsigned)(rm1 >> j) & 1u);
    const unsigned d0 = __ballot_sync(0xffffffffu, (unsigned)(rm0 >> (j + 32)) & 1u), d1 = __ballot_sync(0xffffffffu, (unsigned)(rm1 >> (j + 32)) & 1u);
    if (lane == j) { cm0 = (u64)c0 | ((u64)c1 << 32); cm1 = (u64)d0 | ((u64)d1 << 32); }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = lane + 32 * h;
    if (r < N) {
      const u64 rmask = h ? rm1 : rm0, cmask = h ? cm1 : cm0;
      const size_t o = (size_t)b * N + r;
      rowmask[o] = rmask;
      colmask[o] = cmask;
      nbr[o] = build_record(&rmask, 1, N, !nonbin);
      cnbr[o] = build_record(&cmask, 1, N, !nonbin);
    }
  }
  if (lane == 0) molflag[b] = nonbin ? 0u : 1u;
}

// ---------------------------------------------------------------------------------------------------------
// small batched transposes of the weight matrices (W_l^T for dS = dZ.W^T, Wd^T for the readout GEMM)
// ---------------------------------------------------------------------------------------------------------
struct TransposeJobs {
  const float* in[OCGCN_MAX_LAYERS + 1];
  float* out[OCGCN_MAX_LAYERS + 1];
  int R[OCGCN_MAX_LAYERS + 1];
  int C[OCGCN_MAX_LAYERS + 1];
};
__global__ void transpose_kernel(const TransposeJobs jobs) {
  const int jb = blockIdx.y;
  const float* in = jobs.in[jb];
  float* out = jobs.out[jb];
  const int R = jobs.R[jb], C = jobs.C[jb];
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < R * C; idx += gridDim.x * blockDim.x) {
    const int c = idx / R, r = idx - c * R;  // out is [C][R]
    out[idx] = __ldg(&in[(size_t)r * C + c]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Deterministic column reductions over per-tile / per-split partial rows, one level, no atomics, no fences:
// a CTA of 1024 threads owns COLS columns; thread = (column, row lane), RL = 1024 / COLS row lanes. Every thread folds
// rows lane, lane + RL, ... in ascending order in fp64 with the loads of four rows in flight, the RL lane results are
// then folded in lane order (two fixed-shape steps through shared memory) and one thread per column runs the
// finaliser. The summation order depends only on the shape -> bit-reproducible.
// ---------------------------------------------------------------------------------------------------------
constexpr int kRedThreads = 1024;
// shared-memory doubles wide_reduce<NV, COLS> needs
template <int NV, int COLS>
__host__ __device__ constexpr int wide_reduce_smem() { return NV * kRedThreads + NV * 16 * COLS; }
template <int NV, int COLS, typename Job>
__device__ __forceinline__ void wide_reduce(const Job& job, int P, int cnt, int cgroup, double* sbuf) {
  constexpr int RL = kRedThreads / COLS;
  static_assert(RL >= 16 && RL % 16 == 0, "row lanes are folded 16 groups at a time");
  double (*sred)[RL][COLS] = reinterpret_cast<double (*)[RL][COLS]>(sbuf);
  double (*sred2)[16][COLS] = reinterpret_cast<double (*)[16][COLS]>(sbuf + NV * kRedThreads);
  const int col = threadIdx.x % COLS, rl = threadIdx.x / COLS;
  const int c = cgroup * COLS + col;
  const bool valid = c < cnt;
  double acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.0;
  if (valid) {
    for (int p = rl; p < P; p += RL * 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (p + u * RL < P) job.add(p + u * RL, c, acc);
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) sred[v][rl][col] = acc[v];
  __syncthreads();
  if (rl < 16) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < RL / 16; ++k) t += sred[v][rl * (RL / 16) + k][col];
      sred2[v][rl][col] = t;
    }
  }
  __syncthreads();
  if (rl == 0 && valid) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) t += sred2[v][k][col];
      acc[v] = t;
    }
    job.finish(c, acc);
  }
}

// ---------------------------------------------------------------------------------------------------------
// BatchNorm statistics: merge per-tile (mean, M2) partials -> mean, invstd, scale, shift; running-stat update with
// the unbiased variance (torch.nn.BatchNorm1d semantics, gcn.py:25,40). Sums n_t*mean_t and M2_t + n_t*mean_t^2 in
// fp64 (the tile partials themselves are centred, so the fp64 recombination loses nothing that fp32 inputs hold).
// ---------------------------------------------------------------------------------------------------------
struct BnFinalizeArgs {
  const float* partials;  // [T][2][F]
  int T, F, B, N, MT;
  int training;
  float eps, momentum;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  long long* nbt;
  float* mean;
  float* invstd;
  float* scale;
  float* shift;
};
struct BnJob {
  const BnFinalizeArgs& a;
  __device__ __forceinline__ void add(int t, int c, double* acc) const {
    const double m = (double)__ldg(&a.partials[((size_t)t * 2 + 0) * a.F + c]);
    const double m2 = (double)__ldg(&a.partials[((size_t)t * 2 + 1) * a.F + c]);
    const double nt = (double)(min(a.MT, a.B - t * a.MT) * a.N);
    acc[0] += nt * m;
    acc[1] += m2 + nt * m * m;
  }
  __device__ __forceinline__ void finish(int c, const double* acc) const {
    const double n = (double)a.B * (double)a.N;
    const double mean = acc[0] / n;
    double m2 = acc[1] - n * mean * mean;
    m2 = m2 > 0.0 ? m2 : 0.0;
    const double var = m2 / n;
    const double unbiased = m2 / (n > 1.0 ? n - 1.0 : 1.0);
    a.running_mean[c] = (float)((1.0 - (double)a.momentum) * (double)a.running_mean[c] + (double)a.momentum * mean);
    a.running_var[c] = (float)((1.0 - (double)a.momentum) * (double)a.running_var[c] + (double)a.momentum * unbiased);
    if (c == 0 && a.nbt) *a.nbt += 1;
    const float invstd = (float)(1.0 / sqrt(var + (double)a.eps));
    const float sc = a.gamma[c] * invstd;
    a.mean[c] = (float)mean;
    a.invstd[c] = invstd;
    a.scale[c] = sc;
    a.shift[c] = a.beta[c] - (float)mean * sc;
  }
};
constexpr int kChanCols = 8;      // per-channel jobs: 8 channels x 128 row lanes per CTA
constexpr int kChanColsWide = 2;  // ... or 2 channels x 512 row lanes when there are many partial rows (the job is a latency chain:
                                  // more CTAs and fewer rows per thread shorten it)
static inline bool chan_wide(int rows) { return rows >= 1024; }
template <int COLS>
__global__ void __launch_bounds__(kRedThreads) bn_finalize_kernel(const BnFinalizeArgs a) {
  pdl_wait();
  if (!a.training) {   // eval: running statistics
    const int c = blockIdx.x * COLS + (threadIdx.x % COLS);
    if (threadIdx.x >= COLS || c >= a.F) return;
    const double mean = (double)a.running_mean[c], var = (double)a.running_var[c];
    const float invstd = (float)(1.0 / sqrt(var + (double)a.eps));
    const float sc = a.gamma[c] * invstd;
    a.mean[c] = (float)mean;
    a.invstd[c] = invstd;
    a.scale[c] = sc;
    a.shift[c] = a.beta[c] - (float)mean * sc;
    return;
  }
  __shared__ double sbuf[wide_reduce_smem<2, COLS>()];
  BnJob job{a};
  wide_reduce<2, COLS>(job, a.T, a.F, blockIdx.x, sbuf);
}

// y = scale*z + shift (standalone GraphConvolution output, gcn.py:40)
__global__ void bn_apply_kernel(const float* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                                float* __restrict__ y, size_t total, int F) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % F);
    y[idx] = fmaf(z[idx], __ldg(&scale[c]), __ldg(&shift[c]));
  }
}

// ---------------------------------------------------------------------------------------------------------
// Backward, first half of a layer: dT = pool-scatter(dH) through the saved arg-max, dY = dT*(1-T^2),
// per-tile partial sums of dY and dY*Yhat (the two global sums BatchNorm backward needs).
// plain != 0: the incoming gradient already is dY (standalone GraphConvolution): only the sums are formed.
// dY is written in place over dH.
// ---------------------------------------------------------------------------------------------------------
struct BwdDyArgs {
  int B, N, NW, MT, F, plain;
  const float* adj;
  long long as0, as1, as2;
  const u64* colmask;
  const uint4* cnbr;
  const float* Z;
  const unsigned char* arg;
  const float* mean;
  const float* invstd;
  const float* scale;
  const float* shift;
  float* dH;        // in: dH (or dY when plain), out: dY
  float* partials;  // [tile][2][F]
};
template <int TR>
__global__ void __launch_bounds__(kThreads, 2) bwd_dy_kernel(const BwdDyArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.N, NW = a.NW, F = a.F;
  uint4* sCRec = reinterpret_cast<uint4*>(smem_raw);
  u64* sCol = reinterpret_cast<u64*>(sCRec + TR);
  float* sDH = reinterpret_cast<float*>(sCol + (size_t)TR * NW);
  unsigned char* sArg = reinterpret_cast<unsigned char*>(sDH + TR * KCP);
  float* sWarp = reinterpret_cast<float*>(sArg + TR * KC);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int mol0 = tile * a.MT;
  const int nm = min(a.MT, a.B - mol0);
  const int rows = nm * N;
  const size_t row0 = (size_t)mol0 * N;
  const bool vec2 = (F % 2 == 0);
  if (!a.plain) {
    for (int r = tid; r < rows; r += kThreads) {
      const int m = r / N;
      sCRec[r] = rec_stage(__ldg(&a.cnbr[row0 + r]), m, m * N);
    }
    for (int idx = tid; idx < rows * NW; idx += kThreads) sCol[idx] = a.colmask[row0 * NW + idx];
  }
  __syncthreads();

  constexpr int RPW = TR / kWarps;
  for (int k0 = 0; k0 < F; k0 += KC) {
    const int c = k0 + 2 * lane;
    const bool v0 = c < F, v1 = c + 1 < F;
    float2 mu = make_float2(0.f, 0.f), is = mu, sc = mu, sh = mu;
    if (v0) { mu.x = __ldg(&a.mean[c]); is.x = __ldg(&a.invstd[c]); sc.x = __ldg(&a.scale[c]); sh.x = __ldg(&a.shift[c]); }
    if (v1) { mu.y = __ldg(&a.mean[c + 1]); is.y = __ldg(&a.invstd[c + 1]); sc.y = __ldg(&a.scale[c + 1]); sh.y = __ldg(&a.shift[c + 1]); }
    if (!a.plain) {
#pragma unroll
      for (int h = 0; h < RPW / 8; ++h) {
        float2 v[8];
        uchar2 ag[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int r = warp + (h * 8 + q) * kWarps;
          v[q] = make_float2(0.f, 0.f);
          ag[q] = make_uchar2(0, 0);
          if (r < rows && v0) {
            const size_t o = (row0 + r) * F + c;
            if (vec2) {
              v[q] = *reinterpret_cast<const float2*>(&a.dH[o]);
              ag[q] = __ldg(reinterpret_cast<const uchar2*>(&a.arg[o]));
            } else {
              v[q].x = a.dH[o];
              ag[q].x = __ldg(&a.arg[o]);
              if (v1) { v[q].y = a.dH[o + 1]; ag[q].y = __ldg(&a.arg[o + 1]); }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int r = warp + (h * 8 + q) * kWarps;
          if (r < rows) {
            *reinterpret_cast<float2*>(&sDH[r * KCP + 2 * lane]) = v[q];
            *reinterpret_cast<uchar2*>(&sArg[r * KC + 2 * lane]) = ag[q];
          }
        }
      }
      __syncthreads();
    }
    float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < RPW / 8; ++h) {
      float2 zv[8], gv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = warp + (h * 8 + q) * kWarps;
        const bool ok = r < rows && v0;
        zv[q] = make_float2(0.f, 0.f);
        gv[q] = make_float2(0.f, 0.f);
        if (ok) {
          const size_t o = (row0 + r) * F + c;
          if (vec2) {
            zv[q] = __ldg(reinterpret_cast<const float2*>(&a.Z[o]));
            if (a.plain) gv[q] = *reinterpret_cast<const float2*>(&a.dH[o]);
          } else {
            zv[q].x = __ldg(&a.Z[o]);
            if (v1) zv[q].y = __ldg(&a.Z[o + 1]);
            if (a.plain) { gv[q].x = a.dH[o]; if (v1) gv[q].y = a.dH[o + 1]; }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = warp + (h * 8 + q) * kWarps;
        if (r >= rows || !v0) continue;
        const size_t o = (row0 + r) * F + c;
        const float2 yhat = make_float2((zv[q].x - mu.x) * is.x, (zv[q].y - mu.y) * is.y);
        float2 dy;
        if (a.plain) {
          dy = gv[q];
        } else {
          const uint4 crec = sCRec[r];
          const int m = rec_m(crec), mb = m * N;
          const unsigned char jrel = (unsigned char)(r - mb);
          float2 dt = make_float2(0.f, 0.f);
          if (rec_binary(crec) && !rec_fallback(crec)) {
            // dT[j,c] = sum over rows i that list j as a neighbour and whose arg-max in channel c is j
            const int cnt = rec_cnt(crec);
#define OCGCN_SCAT_STEP(Q)                                                                   \
            if (cnt > Q) {                                                                   \
              const int ir = rec_idx<Q>(crec);                                               \
              const float2 g = *reinterpret_cast<const float2*>(&sDH[ir * KCP + 2 * lane]);  \
              const uchar2 ag = *reinterpret_cast<const uchar2*>(&sArg[ir * KC + 2 * lane]); \
              if (ag.x == jrel) dt.x += g.x;                                                 \
              if (ag.y == jrel) dt.y += g.y;                                                 \
            }
            OCGCN_SCAT_STEP(0) OCGCN_SCAT_STEP(1) OCGCN_SCAT_STEP(2) OCGCN_SCAT_STEP(3) OCGCN_SCAT_STEP(4)
            if (cnt > 5) {
              OCGCN_SCAT_STEP(5) OCGCN_SCAT_STEP(6) OCGCN_SCAT_STEP(7) OCGCN_SCAT_STEP(8)
              if (cnt > 9) { OCGCN_SCAT_STEP(9) OCGCN_SCAT_STEP(10) OCGCN_SCAT_STEP(11) OCGCN_SCAT_STEP(12) }
            }
#undef OCGCN_SCAT_STEP
          } else if (rec_binary(crec)) {
            const u64* mk = &sCol[(size_t)r * NW];
            for (int w = 0; w < NW; ++w) {
              u64 bits = mk[w];
              while (bits) {
                const int ir = mb + 64 * w + __ffsll((long long)bits) - 1;
                bits &= bits - 1;
                if (sArg[ir * KC + 2 * lane] == jrel) dt.x += sDH[ir * KCP + 2 * lane];
                if (sArg[ir * KC + 2 * lane + 1] == jrel) dt.y += sDH[ir * KCP + 2 * lane + 1];
              }
            }
          } else {
            const float* acol = a.adj + (long long)(mol0 + m) * a.as0 + (long long)(r - mb) * a.as2;
            for (int i = 0; i < N; ++i) {
              const float ai = __ldg(&acol[(long long)i * a.as1]);
              if (sArg[(mb + i) * KC + 2 * lane] == jrel) dt.x = fmaf(ai, sDH[(mb + i) * KCP + 2 * lane], dt.x);
              if (sArg[(mb + i) * KC + 2 * lane + 1] == jrel) dt.y = fmaf(ai, sDH[(mb + i) * KCP + 2 * lane + 1], dt.y);
            }
          }
          const float t0 = fast_tanh(fmaf(zv[q].x, sc.x, sh.x)), t1 = fast_tanh(fmaf(zv[q].y, sc.y, sh.y));
          dy = make_float2(dt.x * (1.f - t0 * t0), v1 ? dt.y * (1.f - t1 * t1) : 0.f);
          if (vec2) {
            *reinterpret_cast<float2*>(&a.dH[o]) = dy;
          } else {
            a.dH[o] = dy.x;
            if (v1) a.dH[o + 1] = dy.y;
          }
        }
        s1.x += dy.x; s1.y += dy.y;
        s2.x = fmaf(dy.x, yhat.x, s2.x); s2.y = fmaf(dy.y, yhat.y, s2.y);
      }
    }
    *reinterpret_cast<float2*>(&sWarp[(warp * 32 + lane) * 2]) = s1;
    *reinterpret_cast<float2*>(&sWarp[((kWarps + warp) * 32 + lane) * 2]) = s2;
    __syncthreads();
    if (warp == 0 && v0) {
      float2 t1 = make_float2(0.f, 0.f), t2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const float2 p1 = *reinterpret_cast<const float2*>(&sWarp[(w * 32 + lane) * 2]);
        const float2 p2 = *reinterpret_cast<const float2*>(&sWarp[((kWarps + w) * 32 + lane) * 2]);
        t1.x += p1.x; t1.y += p1.y; t2.x += p2.x; t2.y += p2.y;
      }
      a.partials[((size_t)tile * 2 + 0) * F + c] = t1.x;
      a.partials[((size_t)tile * 2 + 1) * F + c] = t2.x;
      if (v1) { a.partials[((size_t)tile * 2 + 0) * F + c + 1] = t1.y; a.partials[((size_t)tile * 2 + 1) * F + c + 1] = t2.y; }
    }
    __syncthreads();
  }
}

template <int TR>
inline size_t bwd_dy_smem_bytes(int NW) {
  return (size_t)TR * 16 + (size_t)TR * NW * sizeof(u64) + (size_t)TR * KCP * 4 + (size_t)TR * KC + (size_t)4 * kWarps * 32 * 4 + 64;
}

// merge the per-tile (sum dY, sum dY*Yhat) partials in fixed order (fp64) -> dgamma, dbeta, and the
// coefficient vectors of dZ = g*(dY - c1 - Yhat*c2)
struct BwdFinalizeArgs {
  const float* partials;  // [T][2][F]
  int T, F, training;
  double n;
  const float* gamma;
  const float* invstd;
  float* dgamma;
  float* dbeta;
  float* g;
  float* c1;
  float* c2;
};
struct BwdFinJob {
  const BwdFinalizeArgs& a;
  __device__ __forceinline__ void add(int t, int c, double* acc) const {
    acc[0] += (double)__ldg(&a.partials[((size_t)t * 2 + 0) * a.F + c]);
    acc[1] += (double)__ldg(&a.partials[((size_t)t * 2 + 1) * a.F + c]);
  }
  __device__ __forceinline__ void finish(int c, const double* acc) const {
    a.dbeta[c] = (float)acc[0];
    a.dgamma[c] = (float)acc[1];
    a.g[c] = a.gamma[c] * a.invstd[c];
    a.c1[c] = a.training ? (float)(acc[0] / a.n) : 0.f;
    a.c2[c] = a.training ? (float)(acc[1] / a.n) : 0.f;
  }
};
template <int COLS>
__global__ void __launch_bounds__(kRedThreads) bwd_finalize_kernel(const BwdFinalizeArgs a) {
  pdl_wait();
  __shared__ double sbuf[wide_reduce_smem<2, COLS>()];
  BwdFinJob job{a};
  wide_reduce<2, COLS>(job, a.T, a.F, blockIdx.x, sbuf);
}

// ---------------------------------------------------------------------------------------------------------
// dW = A^T . B with A [R][M], B [R][Nn] row-major (R = B*N rows): split-K over CTAs, 128x128 output blocks,
// per-split partials reduced afterwards in fixed order (bit-reproducible, no atomics).
// grid = (splits, mblocks*nblocks)
// ---------------------------------------------------------------------------------------------------------
constexpr int GK = 16;  // rows per staged chunk
__global__ void __launch_bounds__(kThreads, 2) gemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int R, int M,
                                                               int Nn, int rows_per_split, float* __restrict__ partial) {
  __shared__ __align__(16) float sA[GK][128];
  __shared__ __align__(16) float sB[GK][128];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int nbn = (Nn + 127) / 128;
  const int mb = (blockIdx.y / nbn) * 128, nb = (blockIdx.y % nbn) * 128;
  const int r_begin = blockIdx.x * rows_per_split;
  const int r_end = min(R, r_begin + rows_per_split);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int r0 = r_begin; r0 < r_end; r0 += GK) {
    for (int idx = tid; idx < GK * 128; idx += kThreads) {
      const int kk = idx >> 7, cc = idx & 127;
      const int rr = r0 + kk;
      sA[kk][cc] = (rr < r_end && mb + cc < M) ? __ldg(&A[(size_t)rr * M + mb + cc]) : 0.f;
      sB[kk][cc] = (rr < r_end && nb + cc < Nn) ? __ldg(&Bm[(size_t)rr * Nn + nb + cc]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sA[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sA[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sB[kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = partial + (size_t)blockIdx.x * M * Nn;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = mb + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int nn = nb + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (nn < Nn) out[(size_t)m * Nn + nn] = acc[i][j];
    }
  }
}

// out[i] = sum_p partial[p][i] (fixed order) for a LIST of jobs in one launch (every dW_l / db_l / dWd / dbd of a backward
// call): "matrix" jobs = few partial rows (per-CTA split-K partials), many outputs -> 32 columns x 32 row lanes per CTA;
// "vector" jobs = many partial rows (per tile), few outputs -> 8 columns x 128 row lanes per CTA.
constexpr int kMaxReduceJobs = 2 * OCGCN_MAX_LAYERS + 2;
struct ReduceJobs {
  const float* partial[kMaxReduceJobs];
  float* out[kMaxReduceJobs];
  int P[kMaxReduceJobs];
  int count[kMaxReduceJobs];
  int tr_cols[kMaxReduceJobs];    // > 0: each partial holds the TRANSPOSE of the output ([tr_cols][count / tr_cols]) — fused readout gradient dWd^T
  int matrix[kMaxReduceJobs];     // 1: matrix job (32 columns per CTA), 0: vector job (8 columns per CTA)
  int first_cta[kMaxReduceJobs + 1];
  int n;
};
struct SumJob {
  const float* p;
  float* out;
  int cnt;
  int tr_cols;
  __device__ __forceinline__ void add(int q, int c, double* acc) const {
    const int src = tr_cols ? (c % tr_cols) * (cnt / tr_cols) + c / tr_cols : c;
    acc[0] += (double)__ldg(&p[(size_t)q * cnt + src]);
  }
  __device__ __forceinline__ void finish(int c, const double* acc) const { out[c] = (float)acc[0]; }
};
// matrix job, 16-byte loads: a CTA owns 128 consecutive outputs (32 threads x float4) x 32 row lanes; fixed-order fp64 folds
// tr_cols > 0: each partial holds the transpose [tr_cols][cnt / tr_cols] of the [cnt / tr_cols][tr_cols] output (fused readout
// gradient dWd^T): the 16-byte loads run along the partial's rows, the (few) results are stored transposed
__device__ __forceinline__ void wide_reduce_vec4(const float* __restrict__ p, float* __restrict__ out, int P, int cnt, int cgroup, double* sbuf,
                                                 int tr_cols) {
  double (*sred4)[128] = reinterpret_cast<double (*)[128]>(sbuf);   // [32][128]
  const int cq = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = cgroup * 128 + cq * 4;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (c < cnt) {
    for (int q = rl; q < P; q += 32 * 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = (q + u * 32 < P) ? __ldg(reinterpret_cast<const float4*>(&p[(size_t)(q + u * 32) * cnt + c])) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) { acc[0] += (double)v[u].x; acc[1] += (double)v[u].y; acc[2] += (double)v[u].z; acc[3] += (double)v[u].w; }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) sred4[rl][cq * 4 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 128 && cgroup * 128 + (int)threadIdx.x < cnt) {
    double t = 0.0;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) t += sred4[k][threadIdx.x];
    const int c = cgroup * 128 + (int)threadIdx.x;          // index in the partial's layout
    const int tr_rows = tr_cols ? cnt / tr_cols : 0;        // partial is [tr_cols][tr_rows] -> out[(c % tr_rows) * tr_cols + c / tr_rows]
    out[tr_cols ? (c % tr_rows) * tr_cols + c / tr_rows : c] = (float)t;
  }
}
static inline bool reduce_job_vec4(int matrix, int tr_cols, int count) { return matrix && count % 4 == 0 && (tr_cols == 0 || (count / tr_cols) % 4 == 0); }

__global__ void __launch_bounds__(kRedThreads) reduce_partials_kernel(const ReduceJobs jobs) {
  pdl_wait();
  int jb = 0;
  while (jb + 1 < jobs.n && (int)blockIdx.x >= jobs.first_cta[jb + 1]) ++jb;
  const int cg = (int)blockIdx.x - jobs.first_cta[jb];
  __shared__ double sbuf[32 * 128];   // one buffer for whichever path this CTA takes
  static_assert(wide_reduce_smem<1, 32>() <= 32 * 128 && wide_reduce_smem<1, kChanCols>() <= 32 * 128, "shared reduction buffer");
  if (jobs.matrix[jb] && jobs.count[jb] % 4 == 0 && (jobs.tr_cols[jb] == 0 || (jobs.count[jb] / jobs.tr_cols[jb]) % 4 == 0)) {
    wide_reduce_vec4(jobs.partial[jb], jobs.out[jb], jobs.P[jb], jobs.count[jb], cg, sbuf, jobs.tr_cols[jb]);
    return;
  }
  SumJob job{jobs.partial[jb], jobs.out[jb], jobs.count[jb], jobs.tr_cols[jb]};
  if (jobs.matrix[jb]) wide_reduce<1, 32>(job, jobs.P[jb], jobs.count[jb], cg, sbuf);
  else wide_reduce<1, kChanCols>(job, jobs.P[jb], jobs.count[jb], cg, sbuf);
}

}  // namespace ocgcn
