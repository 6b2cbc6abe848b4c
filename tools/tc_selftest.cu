// Stand-alone check of the tcgen05 / TMEM / bulk-copy primitives in openchem_b200/csrc/ocgcn_tc.cuh on a B200.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tc_selftest tools/tc_selftest.cu && ./tc_selftest
// Each test multiplies small matrices on the tensor cores through hand-built no-swizzle operand images and compares
// with a double-precision host result.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../openchem_b200/csrc/ocgcn_tc.cuh"

using namespace ocgcn;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e = (x);                                                                       \
    if (e != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);           \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// Test 1/2: K-major operands built in shared memory by the threads (like a tile-kernel producer), split hi/lo,
// PASSES = 1 (hi.hi) or 3 (hi.hi + lo.hi + hi.lo). C[128][NN] = A[128][K] . B[NN][K]^T
// ---------------------------------------------------------------------------------------------------------
template <bool TF32, int NN>
__global__ void __launch_bounds__(128) kmajor_kernel(const float* A, const float* B, float* C, int K, int passes) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int EPC = TF32 ? 4 : 8;              // elements per 16-byte chunk
  constexpr int ES = TF32 ? 4 : 2;               // element bytes
  const int nchunk = K / EPC;
  const uint32_t lboA = 128 * 16 + 16, lboB = NN * 16 + 16;   // padded chunk pitch
  unsigned char* Ahi = smem;
  unsigned char* Alo = Ahi + nchunk * lboA;
  unsigned char* Bhi = Alo + nchunk * lboA;
  unsigned char* Blo = Bhi + nchunk * lboB;
  uint64_t* bar = reinterpret_cast<uint64_t*>(Blo + nchunk * lboB);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int idx = tid; idx < 128 * K; idx += 128) {
    const int r = idx / K, k = idx % K;
    const float x = A[idx];
    const uint32_t off = (k / EPC) * lboA + r * 16 + (k % EPC) * ES;
    if (TF32) {
      float h, l;
      tc::split_tf32(x, h, l);
      *reinterpret_cast<float*>(Ahi + off) = h;
      *reinterpret_cast<float*>(Alo + off) = l;
    } else {
      __nv_bfloat16 h, l;
      tc::split_bf16(x, h, l);
      *reinterpret_cast<__nv_bfloat16*>(Ahi + off) = h;
      *reinterpret_cast<__nv_bfloat16*>(Alo + off) = l;
    }
  }
  for (int idx = tid; idx < NN * K; idx += 128) {
    const int r = idx / K, k = idx % K;
    const float x = B[idx];
    const uint32_t off = (k / EPC) * lboB + r * 16 + (k % EPC) * ES;
    if (TF32) {
      float h, l;
      tc::split_tf32(x, h, l);
      *reinterpret_cast<float*>(Bhi + off) = h;
      *reinterpret_cast<float*>(Blo + off) = l;
    } else {
      __nv_bfloat16 h, l;
      tc::split_bf16(x, h, l);
      *reinterpret_cast<__nv_bfloat16*>(Bhi + off) = h;
      *reinterpret_cast<__nv_bfloat16*>(Blo + off) = l;
    }
  }
  if (tid == 0) {
    tc::mbar_init(bar, 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(slot, NN);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;

  if (tid == 0) {
    constexpr uint32_t idesc = tc::make_idesc(TF32 ? tc::FMT_TF32 : tc::FMT_BF16, 128, NN, false, false);
    uint32_t acc = 0;
    for (int p = 0; p < passes; ++p) {
      const unsigned char* a = (p == 1) ? Alo : Ahi;
      const unsigned char* b = (p == 2) ? Blo : Bhi;
      for (int ks = 0; ks < nchunk / 2; ++ks) {
        const uint64_t da = tc::make_desc(tc::smem_u32(a) + ks * 2 * lboA, lboA, 128);
        const uint64_t db = tc::make_desc(tc::smem_u32(b) + ks * 2 * lboB, lboB, 128);
        tc::mma_ss<TF32>(tmem, da, db, idesc, acc);
        acc = 1;
      }
    }
    tc::mma_commit(bar);
  }
  tc::mbar_wait(bar, 0);
  tc::fence_after_sync();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < NN; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 32; ++i) C[row * NN + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, NN);
}

// ---------------------------------------------------------------------------------------------------------
// Test 3: MN-major operands (source data [k][mn] row-major), bf16 images prepared in GLOBAL memory by a prep
// kernel in the layout image[mnchunk][k][8], fetched with 1-D bulk copies signalling an mbarrier; the A image is
// also written back with a bulk store (round trip check). C[128][128] = At^T . Bt, single bf16 pass.
// ---------------------------------------------------------------------------------------------------------
__global__ void prep_mn_image(const float* X /*[K][128]*/, __nv_bfloat16* img /*[16][K][8]*/, int K) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < K * 128; idx += gridDim.x * blockDim.x) {
    const int k = idx / 128, mn = idx % 128;
    img[((size_t)(mn / 8) * K + k) * 8 + (mn % 8)] = __float2bfloat16_rn(X[idx]);
  }
}
__global__ void __launch_bounds__(128) mnmajor_kernel(const __nv_bfloat16* Aimg, const __nv_bfloat16* Bimg, float* C, __nv_bfloat16* Aback, int K) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t bytes = 16u * K * 16u;  // one image
  unsigned char* sA = smem;
  unsigned char* sB = smem + bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + bytes);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    tc::mbar_init(bar, 1);
    tc::mbar_init(bar2, 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(slot, 128);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    tc::mbar_expect_tx(bar, 2 * bytes);
    tc::bulk_g2s(sA, Aimg, bytes, bar);
    tc::bulk_g2s(sB, Bimg, bytes, bar);
    tc::mbar_wait(bar, 0);
    tc::fence_after_sync();
    constexpr uint32_t idesc = tc::make_idesc(tc::FMT_BF16, 128, 128, true, true);
    const uint32_t sbo = K * 16;  // between mn chunks
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t da = tc::make_desc(tc::smem_u32(sA) + ks * 256, 128, sbo);
      const uint64_t db = tc::make_desc(tc::smem_u32(sB) + ks * 256, 128, sbo);
      tc::mma_ss<false>(tmem, da, db, idesc, ks > 0);
    }
    tc::mma_commit(bar2);
    tc::bulk_s2g(Aback, sA, bytes);
    tc::bulk_commit();
    tc::bulk_wait0();
  }
  tc::mbar_wait(bar2, 0);
  tc::fence_after_sync();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 32; ++i) C[row * 128 + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

static float bf16_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

template <bool TF32, int NN>
static int run_kmajor(int K, int passes, double tol) {
  std::vector<float> A(128 * K), B(NN * K), C(128 * NN);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dC;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dC, C.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0, C.size() * 4));
  constexpr int EPC = TF32 ? 4 : 8;
  const size_t smem = (size_t)(K / EPC) * (2 * (128 * 16 + 16) + 2 * (NN * 16 + 16)) + 64;
  auto kern = kmajor_kernel<TF32, NN>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<1, 128, smem>>>(dA, dB, dC, K, passes);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  double num = 0, den = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < NN; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * (double)B[n * K + k];
      num += (C[m * NN + n] - s) * (C[m * NN + n] - s);
      den += s * s;
    }
  const double err = sqrt(num / den);
  printf("kmajor %s N=%d K=%d passes=%d: rel-L2 err %.3e (tol %.1e) %s\n", TF32 ? "tf32" : "bf16", NN, K, passes, err, tol,
         err < tol ? "OK" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return err < tol ? 0 : 1;
}

static int run_mnmajor(int K) {
  std::vector<float> A(K * 128), B(K * 128), C(128 * 128);
  srand(2);
  for (auto& v : A) v = bf16_round((float)rand() / RAND_MAX * 2.f - 1.f);
  for (auto& v : B) v = bf16_round((float)rand() / RAND_MAX * 2.f - 1.f);
  float *dA, *dB, *dC;
  __nv_bfloat16 *iA, *iB, *iBack;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dC, C.size() * 4));
  CK(cudaMalloc(&iA, A.size() * 2));
  CK(cudaMalloc(&iB, A.size() * 2));
  CK(cudaMalloc(&iBack, A.size() * 2));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(iBack, 0, A.size() * 2));
  prep_mn_image<<<32, 256>>>(dA, iA, K);
  prep_mn_image<<<32, 256>>>(dB, iB, K);
  const size_t smem = (size_t)2 * 16 * K * 16 + 64;
  CK(cudaFuncSetAttribute(mnmajor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mnmajor_kernel<<<1, 128, smem>>>(iA, iB, dC, iBack, K);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<uint16_t> h1(A.size()), h2(A.size());
  CK(cudaMemcpy(h1.data(), iA, A.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h2.data(), iBack, A.size() * 2, cudaMemcpyDeviceToHost));
  const bool same = memcmp(h1.data(), h2.data(), A.size() * 2) == 0;
  double num = 0, den = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[k * 128 + m] * (double)B[k * 128 + n];
      num += (C[m * 128 + n] - s) * (C[m * 128 + n] - s);
      den += s * s;
    }
  const double err = sqrt(num / den);
  const bool ok = err < 1e-5 && same;
  printf("mnmajor bf16 (bulk g2s/s2g) K=%d: rel-L2 err %.3e, bulk store round trip %s -> %s\n", K, err, same ? "identical" : "DIFFERS",
         ok ? "OK" : "FAIL");
  return ok ? 0 : 1;
}

int main() {
  int fails = 0;
  fails += run_kmajor<false, 128>(64, 1, 6e-3);
  fails += run_kmajor<false, 128>(64, 3, 3e-5);
  fails += run_kmajor<false, 128>(128, 3, 3e-5);
  fails += run_kmajor<false, 256>(64, 3, 3e-5);
  fails += run_kmajor<false, 64>(32, 3, 3e-5);
  fails += run_kmajor<true, 128>(64, 1, 1.5e-3);
  fails += run_kmajor<true, 128>(64, 3, 2e-6);
  fails += run_mnmajor(64);
  fails += run_mnmajor(128);
  printf(fails ? "SELFTEST FAILED (%d)\n" : "SELFTEST PASSED\n", fails);
  return fails ? 1 : 0;
}
