// Blackwell (sm_100a) tensor-core primitives used by the tf32/bf16-class kernels: tcgen05.mma with shared-memory
// operand descriptors, TMEM allocation / loads, mbarriers and 1-D bulk (TMA-engine) copies. Inline PTX only.
//
// Operand images. Every MMA operand is staged in shared memory in the canonical NO-SWIZZLE ("interleaved")
// core-matrix layout: a core matrix is 8 rows of 16 bytes, stored as 128 contiguous bytes. For an operand whose
// rows are the M (or N) index and whose 16-byte chunks run along K ("K-major"):
//     byte offset(row, kchunk) = (row % 8) * 16 + (row / 8) * SBO + kchunk * LBO
// and for an operand stored with the M/N index contiguous ("MN-major", i.e. [k][mn] row-major source data):
//     byte offset(mnchunk, k)  = mnchunk * SBO + (k % 8) * 16 + (k / 8) * LBO
// One tcgen05.mma consumes 32 bytes of K per row (16 bf16 or 8 tf32 values). We always use SBO/LBO such that
// consecutive rows (K-major) or consecutive k (MN-major) are 16 bytes apart, so an image is simply
//     image[chunk][line][16 bytes]   with a (padded) chunk pitch.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ocgcn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor (64 bit) -------------------------------------------------------------
// bits [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 |
// [46,48) version = 1 (sm_100) | [61,64) layout type = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// ---- instruction descriptor (32 bit) for kind::f16 (bf16 inputs) and kind::tf32, fp32 accumulate ----------
// [4,6) D format: 1 = f32 | [7,10) A format, [10,13) B format: 1 = bf16, 2 = tf32 | bit 15 A MN-major | bit 16 B MN-major |
// [17,23) N >> 3 | [24,29) M >> 4
enum : uint32_t { FMT_BF16 = 1u, FMT_TF32 = 2u };
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]; issued by ONE thread on behalf of the CTA
template <bool TF32>
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// all previously issued MMAs of this thread arrive on the mbarrier when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads, bulk copies)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------------------------------------
// one full warp allocates `cols` (power of two >= 32) columns; the base address lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// 32 lanes (this warp's quarter: lane field of taddr = 32 * (warp % 4)) x 32 consecutive fp32 columns:
// thread t receives row (lane base + t), columns col .. col+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  // the suspend-time hint lets the hardware park the warp until the phase completes (or the hint expires) instead of returning to
  // the loop every few hundred cycles: the polling instructions of 16 waiting warps otherwise take issue slots from the co-resident CTA
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(0x989680u)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// ---- 1-D bulk copies on the TMA engine (sizes and addresses multiples of 16 bytes) -----------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
// 2-D tiled TMA load (cp.async.bulk.tensor): box of the tensor map at element coordinates (c0 innermost, c1) -> dense
// [box rows][box cols] image in shared memory (no swizzle); out-of-range elements arrive as zeros and count towards the
// mbarrier's transaction bytes. smem_dst must be 128-byte aligned.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- operand splitting ------------------------------------------------------------------------------------
// x = hi + lo with hi, lo representable in bf16 (lo = round(x - hi)); |x - hi - lo| <= 2^-17 |x|
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {  // a -> low half (lower address)
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
// two fp32 -> packed bf16x2 hi and lo words (element 0 in the low half): two packed conversions (round to nearest even,
// identical to the scalar split), the hi halves are re-expanded with one shift / one mask
__device__ __forceinline__ void split2_bf16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}
// four fp32 -> two packed hi words and two packed lo words (8 bytes of an operand image row each)
__device__ __forceinline__ void split4_bf16(const float4& x, uint2& hi, uint2& lo) {
  split2_bf16(x.x, x.y, hi.x, lo.x);
  split2_bf16(x.z, x.w, hi.y, lo.y);
}
// tf32: hi keeps the top 19 bits (the MMA ignores the low 13 mantissa bits), lo = x - hi (exact in fp32)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = x - hi;
}

}  // namespace tc
}  // namespace ocgcn
