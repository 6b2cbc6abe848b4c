// Tensor-core engine, part 2: weight operand images and the weight-gradient GEMM.
//
//  * wprep_kernel turns fp32 weight matrices into bf16 (hi, lo) operand images in the no-swizzle core-matrix layout
//    the tile kernel's tcgen05.mma reads (ocgcn_tc.cuh): pieces [K/64][ceil(N/128)][part] of [8 k-chunks][128 n][8],
//    zero padded, so that the tile kernel fetches a piece with ONE 16 KB bulk copy.
//  * dw_mma_kernel computes dW = S^T . dZ (gcn.py:35 backward) and dWd = dU^T . H_L as a split-K GEMM on tcgen05:
//    both operands are the bf16 tile images the tile kernels already wrote ([tile][chunks][128 rows][8], i.e.
//    MN-major operands with the 128 rows of a tile as the K dimension). One CTA per slice of tiles: a producer
//    thread streams tile images through a ring of shared-memory stages with bulk copies (TMA engine), one thread
//    issues the MMAs, the fp32 accumulators stay in TMEM for the whole slice; per-CTA partial results are merged
//    afterwards in fixed order (reduce_partials_kernel) -> bit-reproducible, no atomics.
#pragma once
#include "ocgcn_common.cuh"
#include "ocgcn_tc.cuh"

namespace ocgcn {

constexpr int kMaxPrepJobs = 2 * OCGCN_MAX_LAYERS + 2;
struct WPrepJobs {
  const float* src[kMaxPrepJobs];
  void* dst[kMaxPrepJobs];
  int K[kMaxPrepJobs];      // reduction extent of the GEMM the image feeds
  int Nn[kMaxPrepJobs];     // output-column extent
  int form[kMaxPrepJobs];   // 0: element(n,k) = src[k*Nn + n] (src is [K][Nn]);  1: element(n,k) = src[n*K + k] (src is [Nn][K])
  int parts[kMaxPrepJobs];  // 2: hi and lo images, 1: hi only
};
// grid = (8 * max pieces, jobs): one CTA writes one 16-byte-chunk row block (k-chunk kc = blockIdx.x % 8) of the hi(/lo) piece pair
// (q, cbi) — four elements per thread, so the kernel is one load latency deep instead of 32
__global__ void __launch_bounds__(256) wprep_kernel(const WPrepJobs jobs) {
  pdl_wait();
  const int jb = blockIdx.y;
  const int K = jobs.K[jb], Nn = jobs.Nn[jb], parts = jobs.parts[jb];
  const int nq = (K + 63) / 64, ncb = (Nn + 127) / 128;
  const int piece = blockIdx.x >> 3, kc = blockIdx.x & 7;
  if (piece >= nq * ncb) return;
  const int q = piece / ncb, cbi = piece % ncb;
  const float* src = jobs.src[jb];
  const int form = jobs.form[jb];
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(jobs.dst[jb]) + (size_t)piece * parts * 8192 + kc * 1024;
  float v[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int idx = threadIdx.x + 256 * u;
    const int n = idx >> 3, ke = idx & 7;
    const int k = q * 64 + kc * 8 + ke, ng = cbi * 128 + n;
    v[u] = (k < K && ng < Nn) ? (form ? __ldg(&src[(size_t)ng * K + k]) : __ldg(&src[(size_t)k * Nn + ng])) : 0.f;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int idx = threadIdx.x + 256 * u;
    __nv_bfloat16 hi, lo;
    tc::split_bf16(v[u], hi, lo);
    dst[idx] = hi;
    if (parts == 2) dst[8192 + idx] = lo;
  }
}
static inline size_t wimg_bytes(int K, int Nn, int parts) { return (size_t)((K + 63) / 64) * ((Nn + 127) / 128) * parts * 16384; }

struct DwArgs {
  const void* Aimg;   // [tiles][chA][128][8] bf16: rows of the output (M = chA*8 padded channels)
  const void* Bimg;   // [tiles][chB][128][8] bf16: columns of the output
  int tiles, chA, chB;
  int M, Nn;          // true output extents
  int tiles_per_cta, stages;
  int rows_per_stage; // 128: one stage = one whole tile image (one bulk copy per operand); 64: half tiles (one 1 KB copy per chunk),
                      // used when a whole-tile stage would leave room for only one stage
  float* partial;     // [gridDim.x][M][Nn]
};
static inline size_t dw_stage_bytes(int chA, int chB, int rows_per_stage) {
  return ((size_t)(chA > 16 ? chA : 16) + (size_t)chB) * 16 * rows_per_stage;
}

__global__ void __launch_bounds__(128) dw_mma_kernel(const DwArgs a) {
  pdl_wait();
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chA = a.chA, chB = a.chB, S = a.stages;
  const int RS = a.rows_per_stage, halves = 128 / RS;
  const uint32_t cpitch = (uint32_t)RS * 16u;   // bytes of one 16-byte-wide channel chunk inside a stage
  const uint32_t aBytes = (uint32_t)chA * cpitch, aAlloc = (uint32_t)(chA > 16 ? chA : 16) * cpitch, bBytes = (uint32_t)chB * cpitch;
  const uint32_t stageBytes = aAlloc + bBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * stageBytes);
  uint64_t* empty = full + S;
  uint64_t* done = empty + S;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int t0 = blockIdx.x * a.tiles_per_cta;
  const int nt = min(a.tiles, t0 + a.tiles_per_cta) - t0;
  const int mblocks = (chA + 15) / 16, ncols = chB * 8;
  float* out = a.partial + (size_t)blockIdx.x * a.M * a.Nn;
  if (nt <= 0) {   // (cannot happen with the host's split choice; keeps the partial well defined)
    for (int i = tid; i < a.M * a.Nn; i += 128) out[i] = 0.f;
    return;
  }
  uint32_t tcols = 32;
  while ((int)tcols < mblocks * ncols) tcols <<= 1;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(done, 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) {
    __syncwarp();
    tc::tmem_alloc(slot, tcols);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;

  if (warp == 0 && lane == 0) {
    // producer: stage i = rows [RS*(i % halves), +RS) of tile t0 + i / halves; one bulk copy per operand image (whole tiles) or
    // per 16-byte channel chunk (half tiles: the chunk's rows are contiguous in the tile image)
    const unsigned char* ga = reinterpret_cast<const unsigned char*>(a.Aimg) + (size_t)t0 * chA * 2048;
    const unsigned char* gb = reinterpret_cast<const unsigned char*>(a.Bimg) + (size_t)t0 * chB * 2048;
    for (int i = 0; i < nt * halves; ++i) {
      const int s = i % S;
      if (i >= S) tc::mbar_wait(&empty[s], (uint32_t)((i / S) - 1) & 1u);
      unsigned char* st = smem + (size_t)s * stageBytes;
      const size_t ta = (size_t)(i / halves) * chA * 2048 + (size_t)(i % halves) * cpitch;
      const size_t tb = (size_t)(i / halves) * chB * 2048 + (size_t)(i % halves) * cpitch;
      tc::mbar_expect_tx(&full[s], aBytes + bBytes);
      if (halves == 1) {
        tc::bulk_g2s(st, ga + ta, aBytes, &full[s]);
        tc::bulk_g2s(st + aAlloc, gb + tb, bBytes, &full[s]);
      } else {
        for (int c = 0; c < chA; ++c) tc::bulk_g2s(st + (size_t)c * cpitch, ga + ta + (size_t)c * 2048, cpitch, &full[s]);
        for (int c = 0; c < chB; ++c) tc::bulk_g2s(st + aAlloc + (size_t)c * cpitch, gb + tb + (size_t)c * 2048, cpitch, &full[s]);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // MMA issuer: D[m][n] += sum over the tile's 128 rows of A[row][m] * B[row][n]  (both operands MN-major)
    const uint32_t idesc = tc::make_idesc(tc::FMT_BF16, 128, ncols, true, true);
    for (int i = 0; i < nt * halves; ++i) {
      const int s = i % S;
      tc::mbar_wait(&full[s], (uint32_t)(i / S) & 1u);
      tc::fence_after_sync();
      const uint32_t sa = tc::smem_u32(smem + (size_t)s * stageBytes), sb = sa + aAlloc;
      for (int mb = 0; mb < mblocks; ++mb)
        for (int ks = 0; ks < RS / 16; ++ks) {
          const uint64_t da = tc::make_desc(sa + (uint32_t)mb * 16u * cpitch + (uint32_t)ks * 256u, 128, cpitch);
          const uint64_t db = tc::make_desc(sb + (uint32_t)ks * 256u, 128, cpitch);
          tc::mma_ss<false>(tmem + (uint32_t)(mb * ncols), da, db, idesc, (i > 0 || ks > 0) ? 1u : 0u);
        }
      tc::mma_commit(&empty[s]);
    }
    tc::mma_commit(done);
  }
  __syncwarp();
  tc::mbar_wait(done, 0);
  tc::fence_after_sync();
  const bool vec4 = (a.Nn % 4 == 0);
  for (int mb = 0; mb < mblocks; ++mb) {
    const int m = mb * 128 + warp * 32 + lane;
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      float v[32];
      tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mb * ncols + c0), v);
      if (m < a.M && c0 < a.Nn) {
        float* dst = out + (size_t)m * a.Nn + c0;
        if (vec4 && c0 + 32 <= a.Nn) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < a.Nn) dst[i] = v[i];
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, tcols);
}

}  // namespace ocgcn
