// The molecule-tile kernel.
//
// One CTA owns a tile of MT molecules (rows = MT*N <= TR). It streams the tile's channels in chunks of 64
// (two adjacent channels per lane): a PRODUCER turns the chunk into the left GEMM operand entirely in shared
// memory (BatchNorm apply + tanh + neighbour max-pool + A.H gather, or the backward element-wise math), the
// chunk is multiplied with the matching 64 rows of the weight matrix, and an EPILOGUE finishes the tile (bias +
// BatchNorm partial statistics, readout, or the A^T gather). Nothing but the layer boundary tensors ever goes
// to HBM.
//
// Two GEMM engines, selected by the template parameter MM:
//   MM == 0  exact-fp32 mode: FFMA into fp32 register accumulators (8x8 per thread).
//   MM == 1  bf16 tensor-core mode (sm_100a): the producer writes the operand as bf16 hi/lo images in the
//            no-swizzle core-matrix layout (ocgcn_tc.cuh), the weight images arrive by 1-D bulk copies (TMA
//            engine), one thread issues tcgen05.mma (4 passes hi.hi + lo.hi + hi.lo + lo.lo in the forward, 3 in the
//            backward), the fp32 accumulator tile lives in TMEM and is read back with tcgen05.ld for the
//            epilogue. The hi image of the operand is also bulk-stored to HBM: it is the (already formatted)
//            operand of the weight-gradient GEMM (ocgcn_dw.cuh).
//
// Sparsity: the adjacency the data layer produces is 0/1 with ~3 neighbours per atom. adj_pack_kernel turns it
// once per forward into 16-byte neighbour records per row (ocgcn_common.cuh); pool and gather walk those
// records (warp-uniform control flow, two channels per lane). Non-binary molecules take a dense float path.
//
// Reference semantics implemented here: openchem/layers/gcn.py:33-42 and
// openchem/modules/encoders/gcn_encoder.py:43-53 (see include/ocgcn.h).
#pragma once
#include "ocgcn_common.cuh"
#include "ocgcn_tc.cuh"

namespace ocgcn {

template <int TR>
__host__ __device__ constexpr int tile_cols() { return TR == 128 ? 128 : 64; }

constexpr int kImgPitch = 128 * 16 + 16;   // MM: bytes between 16-byte channel chunks of the operand image (padded: conflict-free stores)
constexpr int kWPiece = 8 * 128 * 16;      // MM: bytes of one weight image piece: [8 k-chunks][128 n][16 B]
constexpr int kTmemCols = 128;

// dynamic shared memory needed by tile_kernel<TR, TC, ., ., MM>
template <int TR, int TC, int MM, int NT>
__host__ __device__ inline size_t tile_smem_bytes(int NW) {
  size_t recs = (size_t)2 * TR * 16;
  size_t masks = (size_t)2 * TR * NW * sizeof(u64);
  size_t main_a = MM ? (size_t)TR * (TC + 4) * 4 + (size_t)2 * 8 * kImgPitch   // [chunk buffers | pad] = staging tile, then operand images
                     : (size_t)2 * TR * KCP * 4 + (size_t)KC * TC * 4;          // two chunk buffers + weight chunk
  size_t main_b = (size_t)(TR + 1) * (TC + 4) * 4 + (size_t)(TR + 1) * TC;   // epilogue staging tile + arg-max tile (E_DY / E_ATG_DY), dummy rows
  size_t mainsz = main_a > main_b ? main_a : main_b;
  return recs + masks + mainsz + (size_t)TR * 4 /*molecule flags*/ + (size_t)2 * (NT / 32) * 32 * 4 /*warp partials*/ + 64;
}

// ---- neighbour walks over a staged record: values of two adjacent channels (float2) per lane ------------------
// max with first-index-wins ties (ascending neighbour order), returns tile rows of the arg-max
// The first four neighbours are handled WITHOUT branches so that their shared-memory loads are all in flight at
// once (the walk is latency-bound): a missing neighbour q >= cnt re-reads neighbour 0, which can never win a strict '>'
// comparison (max) or is masked out of the sum. `self` is any valid row of the buffer (used when cnt == 0).
template <int STRIDE>
__device__ __forceinline__ void rec_max2(const uint4& rec, const float* colp, int self, float& b0, float& b1, int& a0, int& a1) {
  const int cnt = rec_cnt(rec);
  const int j0 = cnt > 0 ? rec_idx<0>(rec) : self;
  const int j1 = cnt > 1 ? rec_idx<1>(rec) : j0;
  const int j2 = cnt > 2 ? rec_idx<2>(rec) : j0;
  const int j3 = cnt > 3 ? rec_idx<3>(rec) : j0;
  const float2 v0 = *reinterpret_cast<const float2*>(colp + j0 * STRIDE);
  const float2 v1 = *reinterpret_cast<const float2*>(colp + j1 * STRIDE);
  const float2 v2 = *reinterpret_cast<const float2*>(colp + j2 * STRIDE);
  const float2 v3 = *reinterpret_cast<const float2*>(colp + j3 * STRIDE);
  b0 = cnt > 0 ? v0.x : -INFINITY; b1 = cnt > 0 ? v0.y : -INFINITY; a0 = j0; a1 = j0;
  if (v1.x > b0) { b0 = v1.x; a0 = j1; }
  if (v1.y > b1) { b1 = v1.y; a1 = j1; }
  if (v2.x > b0) { b0 = v2.x; a0 = j2; }
  if (v2.y > b1) { b1 = v2.y; a1 = j2; }
  if (v3.x > b0) { b0 = v3.x; a0 = j3; }
  if (v3.y > b1) { b1 = v3.y; a1 = j3; }
  if (cnt == 0) { b0 = -INFINITY; b1 = -INFINITY; }   // (the re-read row would otherwise beat the -inf start value)
#define OCGCN_MAX_STEP(Q)                                                        \
  if (cnt > Q) {                                                                 \
    const int jr = rec_idx<Q>(rec);                                              \
    const float2 v = *reinterpret_cast<const float2*>(colp + jr * STRIDE);       \
    if (v.x > b0) { b0 = v.x; a0 = jr; }                                         \
    if (v.y > b1) { b1 = v.y; a1 = jr; }                                         \
  }
  if (cnt > 4) {
    OCGCN_MAX_STEP(4) OCGCN_MAX_STEP(5) OCGCN_MAX_STEP(6) OCGCN_MAX_STEP(7) OCGCN_MAX_STEP(8)
    if (cnt > 9) { OCGCN_MAX_STEP(9) OCGCN_MAX_STEP(10) OCGCN_MAX_STEP(11) OCGCN_MAX_STEP(12) }
  }
#undef OCGCN_MAX_STEP
}

template <int STRIDE>
__device__ __forceinline__ float2 rec_sum2(const uint4& rec, const float* colp, int self) {
  const int cnt = rec_cnt(rec);
  const int j0 = cnt > 0 ? rec_idx<0>(rec) : self;
  const int j1 = cnt > 1 ? rec_idx<1>(rec) : j0;
  const int j2 = cnt > 2 ? rec_idx<2>(rec) : j0;
  const int j3 = cnt > 3 ? rec_idx<3>(rec) : j0;
  const float2 v0 = *reinterpret_cast<const float2*>(colp + j0 * STRIDE);
  const float2 v1 = *reinterpret_cast<const float2*>(colp + j1 * STRIDE);
  const float2 v2 = *reinterpret_cast<const float2*>(colp + j2 * STRIDE);
  const float2 v3 = *reinterpret_cast<const float2*>(colp + j3 * STRIDE);
  // ascending neighbour order, exactly the additions the branchy walk performs
  float s0 = cnt > 0 ? v0.x : 0.f, s1 = cnt > 0 ? v0.y : 0.f;
  s0 += cnt > 1 ? v1.x : 0.f; s1 += cnt > 1 ? v1.y : 0.f;
  s0 += cnt > 2 ? v2.x : 0.f; s1 += cnt > 2 ? v2.y : 0.f;
  s0 += cnt > 3 ? v3.x : 0.f; s1 += cnt > 3 ? v3.y : 0.f;
#define OCGCN_SUM_STEP(Q)                                                                       \
  if (cnt > Q) {                                                                                \
    const float2 v = *reinterpret_cast<const float2*>(colp + rec_idx<Q>(rec) * STRIDE);         \
    s0 += v.x; s1 += v.y;                                                                       \
  }
  if (cnt > 4) {
    OCGCN_SUM_STEP(4) OCGCN_SUM_STEP(5) OCGCN_SUM_STEP(6) OCGCN_SUM_STEP(7) OCGCN_SUM_STEP(8)
    if (cnt > 9) { OCGCN_SUM_STEP(9) OCGCN_SUM_STEP(10) OCGCN_SUM_STEP(11) OCGCN_SUM_STEP(12) }
  }
#undef OCGCN_SUM_STEP
  return make_float2(s0, s1);
}

// Variants for records staged with rec_stage_padded (tensor-core engine): unused slots read the neutral dummy row
template <int STRIDE>
__device__ __forceinline__ void rec_max2_padded(const uint4& rec, const float* colp, float& b0, float& b1, int& a0, int& a1) {
  const int cnt = rec_cnt(rec);
  const int j0 = rec_idx<0>(rec), j1 = rec_idx<1>(rec), j2 = rec_idx<2>(rec), j3 = rec_idx<3>(rec);
  const float2 v0 = *reinterpret_cast<const float2*>(colp + j0 * STRIDE);
  const float2 v1 = *reinterpret_cast<const float2*>(colp + j1 * STRIDE);
  const float2 v2 = *reinterpret_cast<const float2*>(colp + j2 * STRIDE);
  const float2 v3 = *reinterpret_cast<const float2*>(colp + j3 * STRIDE);
  b0 = v0.x; b1 = v0.y; a0 = j0; a1 = j0;
  if (v1.x > b0) { b0 = v1.x; a0 = j1; }
  if (v1.y > b1) { b1 = v1.y; a1 = j1; }
  if (v2.x > b0) { b0 = v2.x; a0 = j2; }
  if (v2.y > b1) { b1 = v2.y; a1 = j2; }
  if (v3.x > b0) { b0 = v3.x; a0 = j3; }
  if (v3.y > b1) { b1 = v3.y; a1 = j3; }
#define OCGCN_MAX_STEP(Q)                                                        \
  if (cnt > Q) {                                                                 \
    const int jr = rec_idx<Q>(rec);                                              \
    const float2 v = *reinterpret_cast<const float2*>(colp + jr * STRIDE);       \
    if (v.x > b0) { b0 = v.x; a0 = jr; }                                         \
    if (v.y > b1) { b1 = v.y; a1 = jr; }                                         \
  }
  if (cnt > 4) {
    OCGCN_MAX_STEP(4) OCGCN_MAX_STEP(5) OCGCN_MAX_STEP(6) OCGCN_MAX_STEP(7) OCGCN_MAX_STEP(8)
    if (cnt > 9) { OCGCN_MAX_STEP(9) OCGCN_MAX_STEP(10) OCGCN_MAX_STEP(11) OCGCN_MAX_STEP(12) }
  }
#undef OCGCN_MAX_STEP
}
template <int STRIDE>
__device__ __forceinline__ float2 rec_sum2_padded(const uint4& rec, const float* colp) {
  const int cnt = rec_cnt(rec);
  const float2 v0 = *reinterpret_cast<const float2*>(colp + rec_idx<0>(rec) * STRIDE);
  const float2 v1 = *reinterpret_cast<const float2*>(colp + rec_idx<1>(rec) * STRIDE);
  const float2 v2 = *reinterpret_cast<const float2*>(colp + rec_idx<2>(rec) * STRIDE);
  const float2 v3 = *reinterpret_cast<const float2*>(colp + rec_idx<3>(rec) * STRIDE);
  float s0 = ((v0.x + v1.x) + v2.x) + v3.x, s1 = ((v0.y + v1.y) + v2.y) + v3.y;   // ascending neighbour order; dummies add 0
#define OCGCN_SUM_STEP(Q)                                                                       \
  if (cnt > Q) {                                                                                \
    const float2 v = *reinterpret_cast<const float2*>(colp + rec_idx<Q>(rec) * STRIDE);         \
    s0 += v.x; s1 += v.y;                                                                       \
  }
  if (cnt > 4) {
    OCGCN_SUM_STEP(4) OCGCN_SUM_STEP(5) OCGCN_SUM_STEP(6) OCGCN_SUM_STEP(7) OCGCN_SUM_STEP(8)
    if (cnt > 9) { OCGCN_SUM_STEP(9) OCGCN_SUM_STEP(10) OCGCN_SUM_STEP(11) OCGCN_SUM_STEP(12) }
  }
#undef OCGCN_SUM_STEP
  return make_float2(s0, s1);
}

// ---- four adjacent channels (LDS.128) per lane over PADDED records: the walks of the tensor-core engine's fast paths --------
// A warp covers two tile rows at once (half-warp = one row x 64 channels), or one row x 128 channels; `colp` already points at
// this lane's four channels of buffer row 0. Unused neighbour slots of a padded record point at the buffer's neutral dummy row
// (-inf for the max, 0 for the sums), so the first four steps need no count checks. Same candidate / summation order as the
// two-channel walks above (ascending neighbour index, first index wins ties).
template <int STRIDE>
__device__ __forceinline__ void rec_max4_padded(const uint4& rec, const float* colp, float (&b)[4], int (&a)[4]) {
  const int cnt = rec_cnt(rec);
  const int j0 = rec_idx<0>(rec), j1 = rec_idx<1>(rec), j2 = rec_idx<2>(rec), j3 = rec_idx<3>(rec);
  const float4 v0 = *reinterpret_cast<const float4*>(colp + j0 * STRIDE);
  const float4 v1 = *reinterpret_cast<const float4*>(colp + j1 * STRIDE);
  const float4 v2 = *reinterpret_cast<const float4*>(colp + j2 * STRIDE);
  const float4 v3 = *reinterpret_cast<const float4*>(colp + j3 * STRIDE);
  b[0] = v0.x; b[1] = v0.y; b[2] = v0.z; b[3] = v0.w;
  a[0] = a[1] = a[2] = a[3] = j0;
#define OCGCN_MAX4(V, J)                          \
  if (V.x > b[0]) { b[0] = V.x; a[0] = J; }       \
  if (V.y > b[1]) { b[1] = V.y; a[1] = J; }       \
  if (V.z > b[2]) { b[2] = V.z; a[2] = J; }       \
  if (V.w > b[3]) { b[3] = V.w; a[3] = J; }
  OCGCN_MAX4(v1, j1) OCGCN_MAX4(v2, j2) OCGCN_MAX4(v3, j3)
#define OCGCN_MAX_STEP(Q)                                                        \
  if (cnt > Q) {                                                                 \
    const int jr = rec_idx<Q>(rec);                                              \
    const float4 v = *reinterpret_cast<const float4*>(colp + jr * STRIDE);       \
    OCGCN_MAX4(v, jr)                                                            \
  }
  if (cnt > 4) {
    OCGCN_MAX_STEP(4) OCGCN_MAX_STEP(5) OCGCN_MAX_STEP(6) OCGCN_MAX_STEP(7) OCGCN_MAX_STEP(8)
    if (cnt > 9) { OCGCN_MAX_STEP(9) OCGCN_MAX_STEP(10) OCGCN_MAX_STEP(11) OCGCN_MAX_STEP(12) }
  }
#undef OCGCN_MAX_STEP
#undef OCGCN_MAX4
}
template <int STRIDE>
__device__ __forceinline__ float4 rec_sum4_padded(const uint4& rec, const float* colp) {
  const int cnt = rec_cnt(rec);
  const float4 v0 = *reinterpret_cast<const float4*>(colp + rec_idx<0>(rec) * STRIDE);
  const float4 v1 = *reinterpret_cast<const float4*>(colp + rec_idx<1>(rec) * STRIDE);
  const float4 v2 = *reinterpret_cast<const float4*>(colp + rec_idx<2>(rec) * STRIDE);
  const float4 v3 = *reinterpret_cast<const float4*>(colp + rec_idx<3>(rec) * STRIDE);
  float4 s = make_float4(((v0.x + v1.x) + v2.x) + v3.x, ((v0.y + v1.y) + v2.y) + v3.y, ((v0.z + v1.z) + v2.z) + v3.z,
                         ((v0.w + v1.w) + v2.w) + v3.w);   // ascending neighbour order; dummies add 0
#define OCGCN_SUM_STEP(Q)                                                                       \
  if (cnt > Q) {                                                                                \
    const float4 v = *reinterpret_cast<const float4*>(colp + rec_idx<Q>(rec) * STRIDE);         \
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;                                             \
  }
  if (cnt > 4) {
    OCGCN_SUM_STEP(4) OCGCN_SUM_STEP(5) OCGCN_SUM_STEP(6) OCGCN_SUM_STEP(7) OCGCN_SUM_STEP(8)
    if (cnt > 9) { OCGCN_SUM_STEP(9) OCGCN_SUM_STEP(10) OCGCN_SUM_STEP(11) OCGCN_SUM_STEP(12) }
  }
#undef OCGCN_SUM_STEP
  return s;
}

template <int STRIDE>
__device__ __forceinline__ float4 rec_sum4(const uint4& rec, const float* colp) {
  const int cnt = rec_cnt(rec);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#define OCGCN_SUM_STEP(Q)                                                                       \
  if (cnt > Q) {                                                                                \
    const float4 v = *reinterpret_cast<const float4*>(colp + rec_idx<Q>(rec) * STRIDE);         \
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;                                             \
  }
  OCGCN_SUM_STEP(0) OCGCN_SUM_STEP(1) OCGCN_SUM_STEP(2) OCGCN_SUM_STEP(3) OCGCN_SUM_STEP(4)
  if (cnt > 5) {
    OCGCN_SUM_STEP(5) OCGCN_SUM_STEP(6) OCGCN_SUM_STEP(7) OCGCN_SUM_STEP(8)
    if (cnt > 9) { OCGCN_SUM_STEP(9) OCGCN_SUM_STEP(10) OCGCN_SUM_STEP(11) OCGCN_SUM_STEP(12) }
  }
#undef OCGCN_SUM_STEP
  return s;
}

// two adjacent channels of one row from global memory (8-byte access when the row pitch allows it)
__device__ __forceinline__ float2 ldg2(const float* p, bool v0, bool v1, bool vec) {
  if (vec) return v0 ? __ldg(reinterpret_cast<const float2*>(p)) : make_float2(0.f, 0.f);   // even width: v1 == v0
  return make_float2(v0 ? __ldg(p) : 0.f, v1 ? __ldg(p + 1) : 0.f);
}
__device__ __forceinline__ void stg2(float* p, float2 v, bool v0, bool v1, bool vec) {
  if (vec) { if (v0) *reinterpret_cast<float2*>(p) = v; return; }
  if (v0) p[0] = v.x;
  if (v1) p[1] = v.y;
}

template <int TR, int TC, int PROD, int EPI, int MM, int NT>
__global__ void __launch_bounds__(NT, (TR == 128 ? 2 : 1)) tile_kernel(const TileArgs a, const __grid_constant__ CUtensorMap tm_in0) {
  constexpr int NWARP = NT / 32;
  constexpr int TYN = TR / 8;
  constexpr int TXN = TC / 8;
  static_assert(MM == 1 || TYN * TXN == NT, "thread grid must cover the tile");
  static_assert(MM == 0 || (TR == 128 && TC == 128), "the tensor-core path works on 128x128 tiles");
  constexpr int TCP = TC + 4;
  constexpr int RPW = TR / NWARP;      // rows per warp (16 or 32), processed in batches of 8 independent loads
  constexpr int BS = MM ? KC : KCP;     // chunk buffer row stride (floats); the FFMA fragment loads want the padding
  // MM: every operand is split x = hi + lo (bf16 each). Forward products take 4 passes (hi.hi + lo.hi + hi.lo + lo.lo:
  // fp32-class values, so the pool arg-max decisions match an fp32 forward), backward products 3 (no lo.lo).
  constexpr int NPASS = (PROD == P_X || PROD == P_MID || PROD == P_RO) ? 4 : 3;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int N = a.N, NW = a.NW, K = a.K, NC = a.NC;
  // Layout: [records][main region][flags][warp partials][barriers][tmem slot][row masks][column masks]. Only the bit masks have
  // a run-time size (NW words per row); they come LAST so that every other region sits at a compile-time offset and its
  // addresses fold into instruction immediates instead of being re-formed from NW inside the loops.
  uint4* sRec = reinterpret_cast<uint4*>(smem_raw);
  uint4* sCRec = sRec + TR;
  float* mainp = reinterpret_cast<float*>(sCRec + TR);
  constexpr int BROWS = MM ? TR + 1 : TR;   // MM: row TR of each chunk buffer is the neutral dummy row of the padded neighbour records
  float* buf0 = mainp;
  float* buf1 = buf0 + BROWS * BS;
  float* sW = buf1 + TR * BS;                                          // MM == 0: weight chunk
  unsigned char* imgHi = reinterpret_cast<unsigned char*>(mainp + TR * TCP);  // MM == 1: operand images [8][kImgPitch], after the staging tile
  unsigned char* imgLo = imgHi + 8 * kImgPitch;
  float* const zbox0 = reinterpret_cast<float*>(imgHi);   // TMA landing buffer of a tile's first Zprev chunk (see use_tma below)
  static_assert(!MM || (size_t)128 * KC * 4 <= (size_t)2 * 8 * kImgPitch, "the Zprev box fits the operand-image region");
  unsigned char* wHi = reinterpret_cast<unsigned char*>(buf0);                 // MM == 1: weight image pieces (hi, lo) alias buf0
  float* sC = mainp;  // epilogue staging, aliases buf0/buf1/sW
  constexpr size_t main_a = MM ? (size_t)TR * TCP + (size_t)2 * 8 * kImgPitch / 4 : (size_t)2 * TR * KCP + (size_t)KC * TC;
  constexpr size_t main_b = (size_t)(TR + 1) * TCP + (size_t)(TR + 1) * TC / 4;
  constexpr size_t mainsz = main_a > main_b ? main_a : main_b;
  unsigned* sFlag = reinterpret_cast<unsigned*>(mainp + mainsz);
  float* sWarp = reinterpret_cast<float*>(sFlag + TR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sWarp + 2 * NWARP * 32);   // MM: [0] weight image landed, [1] MMAs complete, [2] second column block's weight image landed, [3] see tslot
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);   // ([3]: saved operand image of the fused dW GEMM landed)
  u64* sRow = reinterpret_cast<u64*>(bars + 8);               // (bars + 4 .. + 8: tslot and padding, 64 bytes after the warp partials in all)
  u64* sCol = sRow + (size_t)TR * NW;

  const int tid = threadIdx.x, lane = tid & 31;
  // warp index through a shuffle: the compiler then KNOWS it is warp-uniform, keeps everything derived from it (row index,
  // neighbour record, its flags and counts) on the uniform datapath and emits uniform branches instead of divergence scaffolding
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int tx = tid % TXN, ty = tid / TXN;
  const bool kvec2 = (K % 2 == 0);
  // T goes to the buffer the weight image does NOT alias while the pool still reads it
  float* const bufT = (MM && PROD == P_RO) ? buf1 : buf0;
  float* const bufH = (MM && PROD == P_RO) ? buf0 : buf1;
  // this lane's slot inside a 64-channel operand image row: 16-byte chunk lane/4, word lane%4
  const uint32_t img_lane_off = (uint32_t)(lane >> 2) * kImgPitch + (uint32_t)(lane & 3) * 4;
  // fused weight gradient (backward producers of the tensor-core engine, widths <= 128): dW (+)= saved_image^T . produced_image
  // accumulated in TMEM columns [128, 256) across all tiles of this persistent CTA (ocgcn_api.cu decides via a.dw_img)
  const bool do_dw = MM && (PROD == P_DZ || PROD == P_DU) && a.dw_img != nullptr;
  const bool use_mma = MM && (!a.skip_gemm || do_dw);
  uint32_t tmem = 0, ph_w = 0, ph_w1 = 0, ph_mma = 0, ph_s = 0;
  const bool use_tma = MM && (PROD == P_MID || PROD == P_RO) && a.use_tma;
  constexpr uint32_t kZBoxBytes = 128u * KC * 4u;
  const uint32_t tcols = (MM && (a.ncb > 1 || do_dw)) ? 2 * kTmemCols : kTmemCols;   // <= 2 accumulators of 128 columns
  bool mma_pending = false, store_pending = false;

  if (use_mma) {
    if (tid == 0) {
      tc::mbar_init(&bars[0], 1);
      tc::mbar_init(&bars[1], 1);
      tc::mbar_init(&bars[2], 1);
      tc::mbar_init(&bars[3], 1);
      tc::mbar_init(&bars[5], 1);   // [5]: TMA-staged producer input chunk landed (bars[4] holds the TMEM slot)
      tc::mbar_init_fence();
    }
    if (warp == 0) {
      __syncwarp();
      tc::tmem_alloc(tslot, tcols);
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (use_mma) {
    tc::fence_after_sync();
    tmem = *tslot;
  }
  pdl_wait();   // everything above touched only shared memory / TMEM; the predecessor kernel's results are needed from here on

  // persistent over tiles (MM launches use at most two CTAs per SM; the exact-fp32 engine launches one CTA per tile)
  int tile_iter = 0;
  for (int tix = blockIdx.x; tix < a.ntiles; tix += gridDim.x, ++tile_iter) {
  // a.reverse: walk the batch from its last tile to its first. Consecutive kernels of a pass alternate the direction, so each starts
  // on the tiles its predecessor wrote LAST (the part of Z / dY that is still in the L2)
  const int tile = a.reverse ? a.ntiles - 1 - tix : tix;
  const int mol0 = tile * a.MT;
  const int nm = min(a.MT, a.B - mol0);
  const int rows = nm * N;
  const size_t row0 = (size_t)mol0 * N;
  // tile base pointers: the 64-bit products are formed once per tile, the loops below add 32-bit offsets (rows * width < 2^16)
  const float* const in0_t = (PROD == P_X || !a.in0) ? a.in0 : a.in0 + row0 * K;
  const float* const in1_t = (PROD == P_DZ) ? a.in1 + row0 * K : a.in1;
  unsigned char* const argo_t = a.arg_out ? a.arg_out + row0 * K : nullptr;
  float* const prodo_t = a.prod_out ? a.prod_out + row0 * K : nullptr;
  float* const out0_t = a.out0 ? a.out0 + row0 * NC : nullptr;
  const unsigned char* const argi_t = a.arg_in ? a.arg_in + row0 * NC : nullptr;
  const float* const zprev_t = a.zprev ? a.zprev + row0 * NC : nullptr;
  const float* const o_t = (PROD == P_DU) ? a.in2 + (size_t)mol0 * K : nullptr;
  const float* const do_t = (PROD == P_DU) ? a.in1 + (size_t)mol0 * K : nullptr;

  // First channel chunk of Zprev -> zbox0, the operand-image region (dead between the last MMA of a tile and the first gather of the
  // next one). The first tile of this CTA asks for it here; every later tile's box was requested during the previous tile's
  // epilogue (below), so it has landed by the time the tanh stage wants it.
  if (use_tma && tid == 0 && tile_iter == 0) {
    tc::fence_async_smem();
    tc::mbar_expect_tx(&bars[5], kZBoxBytes);
    tc::tma_load_2d(zbox0, &tm_in0, 0, (int)row0, &bars[5]);
  }
  // ---- stage neighbour records (made tile-relative), masks and per-molecule flags ---------------------------
  int simple = 1;   // every row AND column of the tile is a binary record with <= kMaxNbr neighbours (any molecular graph)
  for (int r = tid; r < TR; r += NT) {
    uint4 rec = make_uint4(0u, 0u, 0u, 0u), crec = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows) {
      const int m = r / N, mb = m * N;
      rec = MM ? rec_stage_padded(__ldg(&a.nbr[row0 + r]), m, mb, TR) : rec_stage(__ldg(&a.nbr[row0 + r]), m, mb);
      crec = MM ? rec_stage_padded(__ldg(&a.cnbr[row0 + r]), m, mb, TR) : rec_stage(__ldg(&a.cnbr[row0 + r]), m, mb);
      simple &= (int)(rec_binary(rec) && !rec_fallback(rec) && rec_binary(crec) && !rec_fallback(crec));
    }
    sRec[r] = rec;
    sCRec[r] = crec;
  }
  for (int idx = tid; idx < rows * NW; idx += NT) {
    sRow[idx] = a.rowmask[row0 * NW + idx];
    sCol[idx] = a.colmask[row0 * NW + idx];
  }
  // tile_simple selects row loops that contain ONLY the padded-record walk and are unrolled over the warp's RPW rows (their
  // record, buffer and image addresses become immediates); hub rows / float adjacency take the general loops
  const bool tile_simple = __syncthreads_and(simple) != 0 && MM;
  // fast paths of the tensor-core engine (any molecular graph, widths that are multiples of four): FOUR adjacent channels per
  // lane (LDS.128 / LDG.128); a half-warp owns one tile row of the 64-channel chunk, so one warp instruction covers two rows
  const bool fast4 = tile_simple && (K % 4 == 0);
  const int hl = lane & 15, hh = lane >> 4;

  float acc[MM ? 1 : 8][MM ? 1 : 8];

  for (int cb = 0; cb < NC || (a.skip_gemm && cb == 0); cb += TC) {
    const bool first_pass = (cb == 0);
    if (!MM) {
#pragma unroll
      for (int r = 0; r < (MM ? 1 : 8); ++r)
#pragma unroll
        for (int c = 0; c < (MM ? 1 : 8); ++c) acc[r][c] = 0.f;
    }

    // MM: the operand images are produced ONCE; the MMAs of every output column block are issued from them into separate
    // TMEM accumulators, and the later column passes only run their epilogue
    for (int k0 = 0; k0 < K && (!MM || cb == 0); k0 += KC) {
      const int c = k0 + 2 * lane;  // this lane's first channel
      const bool v0 = c < K, v1 = c + 1 < K;
      const float* opnd = buf0;

      // MM: the previous chunk's MMAs read the weight image (buf0 / buf1) and the operand images, and a bulk store may still
      // read the hi image. Every producer first ISSUES its global loads of this chunk, then calls chunk_sync() before its
      // first shared-memory write: the MMA tail hides behind the load latency. Thread 0 then starts this chunk's bulk loads.
      auto chunk_sync = [&]() {
        if (!MM) return;
        // ONE warp polls the mbarrier, everybody else parks at the hardware barrier: waiting warps then take no issue slots from the
        // co-resident CTA (16 polling warps were 7-10 % of the executed instructions)
        bool resync = false;
        if (use_tma) {   // this chunk's TMA box has landed in buf1
          // (one box per chunk since the kernel started: the phase parity follows from the chunk count, no register carried)
          if (warp == 0) tc::mbar_wait(&bars[5], (uint32_t)(tile_iter * ((K + KC - 1) / KC) + k0 / KC) & 1u);
          resync = true;
        }
        if (mma_pending) {
          if (warp == 0) tc::mbar_wait(&bars[1], ph_mma);
          ph_mma ^= 1;
          mma_pending = false;
          resync = true;
        }
        // producer-only launches (skip_gemm) need no lo image: odd chunks use its space as a second hi image, so the
        // bulk store of one chunk overlaps the production of the next and the wait is only needed every other chunk
        const bool reuse = !a.skip_gemm || do_dw || (k0 / KC) >= 2;
        if (store_pending && reuse) {
          if (tid == 0) tc::bulk_wait_read0();   // the previous chunks' images have been read by the bulk stores
          if (PROD == P_DZ || PROD == P_DU) resync = true;   // these producers write the image in their first stage
          store_pending = false;
        }
        if (resync) __syncthreads();
        if (do_dw && k0 == 0 && tid == 0) {   // this tile's saved operand image (S_l or H_L) -> buf1 (free in these producers)
          tc::mbar_expect_tx(&bars[3], (uint32_t)a.dw_chunks * 2048u);
          tc::bulk_g2s(buf1, reinterpret_cast<const unsigned char*>(a.dw_img) + (size_t)tile * a.dw_chunks * 2048, (uint32_t)a.dw_chunks * 2048u, &bars[3]);
        }
        if (use_mma && !a.skip_gemm && tid == 0 && !(PROD == P_MID)) {   // buf0 is free for the whole chunk: fetch the weight pieces now
          const unsigned char* src = reinterpret_cast<const unsigned char*>(a.wimg) +
                                     ((size_t)(k0 / KC) * a.ncb + (cb / TC)) * 2 * kWPiece;
          tc::mbar_expect_tx(&bars[0], 2 * kWPiece);
          tc::bulk_g2s(wHi, src, kWPiece, &bars[0]);
          tc::bulk_g2s(wHi + kWPiece, src + kWPiece, kWPiece, &bars[0]);
        }
      };
      if (MM) {
      } else if (!a.skip_gemm) {
        // weight chunk [KC][TC] (independent of the producer; sW is free after the previous trailing sync)
        if (NC % 4 == 0) {
          for (int idx = tid; idx < KC * TC / 4; idx += NT) {
            const int kk = idx / (TC / 4), cc = (idx - kk * (TC / 4)) * 4;
            const int gk = k0 + kk, gc = cb + cc;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gk < K && gc < NC) w = __ldg(reinterpret_cast<const float4*>(&a.Bm[(size_t)gk * NC + gc]));
            *reinterpret_cast<float4*>(&sW[kk * TC + cc]) = w;
          }
        } else {
          for (int idx = tid; idx < KC * TC; idx += NT) {
            const int kk = idx / TC, cc = idx - kk * TC;
            const int gk = k0 + kk, gc = cb + cc;
            sW[idx] = (gk < K && gc < NC) ? __ldg(&a.Bm[(size_t)gk * NC + gc]) : 0.f;
          }
        }
      }

      // the producer's result for (row r, channels c, c+1): fp32 chunk buffer (FFMA) or bf16 hi/lo operand images (MMA)
      auto put_operand = [&](float* dst, int r, float2 val) {
        if (MM) {
          uint32_t hi, lo;
          tc::split2_bf16(val.x, val.y, hi, lo);
          const uint32_t off = img_lane_off + (uint32_t)r * 16;
          if (a.skip_gemm && !do_dw) {
            *reinterpret_cast<uint32_t*>(((k0 / KC) & 1 ? imgLo : imgHi) + off) = hi;
          } else {
            *reinterpret_cast<uint32_t*>(imgHi + off) = hi;
            *reinterpret_cast<uint32_t*>(imgLo + off) = lo;
          }
        } else {
          *reinterpret_cast<float2*>(&dst[r * BS + 2 * lane]) = val;
        }
      };

      // fast4: this lane's four channels of row r -> 8 bytes of the hi image and 8 bytes of the lo image
      const uint32_t img_lane_off4 = (uint32_t)(hl >> 1) * kImgPitch + (uint32_t)(hl & 1) * 8;
      auto put_operand4 = [&](int r, const float4& val) {
        uint2 hi, lo;
        tc::split4_bf16(val, hi, lo);
        const uint32_t off = img_lane_off4 + (uint32_t)r * 16;
        if (a.skip_gemm && !do_dw) {
          *reinterpret_cast<uint2*>(((k0 / KC) & 1 ? imgLo : imgHi) + off) = hi;
        } else {
          *reinterpret_cast<uint2*>(imgHi + off) = hi;
          *reinterpret_cast<uint2*>(imgLo + off) = lo;
        }
      };
      const int c4 = k0 + 4 * hl;   // fast4: this lane's first channel
      const bool cv = c4 < K;       // (K % 4 == 0: all four channels valid, or none)

      // =========================== PRODUCER ===========================
      if (PROD == P_X) {
        if (a.x_bits) {
          // compact batch: the row's bit word now, shift + convert after the wait (a use inside the loop would serialise the loads)
#pragma unroll
          for (int h = 0; h < RPW / 8; ++h) {
            uint32_t wv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int r = warp + (h * 8 + q) * NWARP;
              wv[q] = 0u;
              if (r < rows && v0) {
                const int m = rec_m(sRec[r]), i = r - m * N;
                wv[q] = __ldg(&a.x_bits[((size_t)(mol0 + m) * N + i) * a.xw + (c >> 5)]);
              }
            }
            if (h == 0) chunk_sync();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint32_t w = wv[q] >> (c & 31);     // c is even: channels c and c + 1 share the word
              *reinterpret_cast<float2*>(&buf1[(warp + (h * 8 + q) * NWARP) * BS + 2 * lane]) =
                  make_float2((float)(w & 1u), v1 ? (float)((w >> 1) & 1u) : 0.f);
            }
          }
        } else {
#pragma unroll
        for (int h = 0; h < RPW / 8; ++h) {
          float2 v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int r = warp + (h * 8 + q) * NWARP;
            v[q] = make_float2(0.f, 0.f);
            if (r < rows && v0) {
              const int m = rec_m(sRec[r]), i = r - m * N;
              const float* px = a.in0 + (long long)(mol0 + m) * a.xs0 + (long long)i * a.xs1 + (long long)c * a.xs2;
              v[q].x = __ldg(px);
              if (v1) v[q].y = __ldg(px + a.xs2);
            }
          }
          if (h == 0) chunk_sync();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float2*>(&buf1[(warp + (h * 8 + q) * NWARP) * BS + 2 * lane]) = v[q];
        }
        }
        if (MM && warp == 0) *reinterpret_cast<float2*>(&buf1[TR * BS + 2 * lane]) = make_float2(0.f, 0.f);   // dummy row of the padded records
        __syncthreads();
      } else if (PROD == P_MID || PROD == P_RO) {
        float2 sc = make_float2(0.f, 0.f), sh = make_float2(0.f, 0.f);
        if (v0) { sc.x = __ldg(&a.vec0[c]); sh.x = __ldg(&a.vec1[c]); }
        if (v1) { sc.y = __ldg(&a.vec0[c + 1]); sh.y = __ldg(&a.vec1[c + 1]); }
        // (the 2*log2(e) factor of the exp2-based tanh is folded into the lane's scale / shift: one multiply less per element)
        const float2 scs = make_float2(sc.x * kTanhScale, sc.y * kTanhScale), shs = make_float2(sh.x * kTanhScale, sh.y * kTanhScale);
        if (MM && fast4) {
          float4 sc4 = make_float4(0.f, 0.f, 0.f, 0.f), sh4 = sc4;
          if (cv) { sc4 = __ldg(reinterpret_cast<const float4*>(&a.vec0[c4])); sh4 = __ldg(reinterpret_cast<const float4*>(&a.vec1[c4])); }
          sc4.x *= kTanhScale; sc4.y *= kTanhScale; sc4.z *= kTanhScale; sc4.w *= kTanhScale;
          sh4.x *= kTanhScale; sh4.y *= kTanhScale; sh4.z *= kTanhScale; sh4.w *= kTanhScale;
          float4 v[RPW / 2];
          if (!use_tma) {
#pragma unroll
            for (int sI = 0; sI < RPW / 2; ++sI) {
              const int r = warp + (2 * sI + hh) * NWARP;
              v[sI] = (r < rows && cv) ? __ldg(reinterpret_cast<const float4*>(&in0_t[r * K + c4])) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          chunk_sync();
          if (use_tma) {   // the chunk sits in shared memory as a dense [128][64] box: chunk 0 in zbox0, later chunks in buf1 (P_RO: tanh in place there)
            const float* const zsrc = (k0 == 0) ? zbox0 : buf1;
#pragma unroll
            for (int sI = 0; sI < RPW / 2; ++sI) v[sI] = *reinterpret_cast<const float4*>(&zsrc[(warp + (2 * sI + hh) * NWARP) * BS + 4 * hl]);
          }
#pragma unroll
          for (int sI = 0; sI < RPW / 2; ++sI) {
            const int r = warp + (2 * sI + hh) * NWARP;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows && cv) {
              t.x = fast_tanh_scaled(fmaf(v[sI].x, sc4.x, sh4.x));
              t.y = fast_tanh_scaled(fmaf(v[sI].y, sc4.y, sh4.y));
              t.z = fast_tanh_scaled(fmaf(v[sI].z, sc4.z, sh4.z));
              t.w = fast_tanh_scaled(fmaf(v[sI].w, sc4.w, sh4.w));
            }
            *reinterpret_cast<float4*>(&bufT[r * BS + 4 * hl]) = t;
          }
        } else {
          constexpr int TB = (RPW > 16) ? 16 : RPW;   // rows per batch: all their loads are in flight before the first tanh
#pragma unroll
          for (int h = 0; h < RPW / TB; ++h) {
            float2 v[TB];
#pragma unroll
            for (int q = 0; q < TB; ++q) {
              const int r = warp + (h * TB + q) * NWARP;
              v[q] = (r < rows && v0) ? ldg2(&in0_t[r * K + c], v0, v1, kvec2) : make_float2(0.f, 0.f);
            }
            if (h == 0) chunk_sync();
#pragma unroll
            for (int q = 0; q < TB; ++q) {
              const int r = warp + (h * TB + q) * NWARP;
              float2 t = make_float2(0.f, 0.f);
              if (r < rows) {
                if (v0) t.x = fast_tanh_scaled(fmaf(v[q].x, scs.x, shs.x));
                if (v1) t.y = fast_tanh_scaled(fmaf(v[q].y, scs.y, shs.y));
              }
              *reinterpret_cast<float2*>(&bufT[r * BS + 2 * lane]) = t;
            }
          }
        }
        if (MM && warp == 0) *reinterpret_cast<float2*>(&bufT[TR * BS + 2 * lane]) = make_float2(-INFINITY, -INFINITY);   // dummy row: never wins a max
        __syncthreads();
        // neighbour max-pool: P[i,c] = max_j A[i,j]*T[j,c], first index wins ties (gcn_encoder.py:44-50)
        if (MM && fast4) {
          // (readout producer: the operand split and image stores sit inside this loop, two row pairs in flight fit the 64 registers)
#pragma unroll(PROD == P_RO ? 2 : RPW / 2)
          for (int sI = 0; sI < RPW / 2; ++sI) {
            const int r = warp + (2 * sI + hh) * NWARP;
            float bq[4] = {0.f, 0.f, 0.f, 0.f};
            if (r < rows) {
              const uint4 rec = sRec[r];
              const int mb = rec_m(rec) * N;
              int aq[4];
              rec_max4_padded<BS>(rec, &bufT[4 * hl], bq, aq);
              if (rec_has_zero(rec)) {  // a zero entry A[i,jz] contributes the candidate 0*T = 0 (first index wins an exact tie)
                const int jz = mb + rec_jz(rec);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const bool keep = bq[e] > 0.f || (bq[e] == 0.f && aq[e] < jz);
                  aq[e] = keep ? aq[e] : jz;
                  bq[e] = fmaxf(bq[e], 0.f);
                }
              }
              if (first_pass && cv)
                *reinterpret_cast<uchar4*>(argo_t + (r * K + c4)) =
                    make_uchar4((unsigned char)(aq[0] - mb), (unsigned char)(aq[1] - mb), (unsigned char)(aq[2] - mb), (unsigned char)(aq[3] - mb));
            }
            const float4 h4 = cv ? make_float4(bq[0], bq[1], bq[2], bq[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (PROD == P_RO) put_operand4(r, h4);
            else *reinterpret_cast<float4*>(&bufH[r * BS + 4 * hl]) = h4;
          }
        } else if (tile_simple) {
#pragma unroll 4
          for (int q = 0; q < RPW; ++q) {
            const int r = warp + q * NWARP;
            float b0 = 0.f, b1 = 0.f;
            if (r < rows) {
              const uint4 rec = sRec[r];
              const int mb = rec_m(rec) * N;
              int a0 = mb, a1 = mb;
              rec_max2_padded<BS>(rec, &bufT[2 * lane], b0, b1, a0, a1);
              if (rec_has_zero(rec)) {  // a zero entry A[i,jz] contributes the candidate 0*T = 0
                const int jz = mb + rec_jz(rec);
                if (!(b0 > 0.f)) { if (b0 < 0.f || jz < a0) a0 = jz; b0 = 0.f; }
                if (!(b1 > 0.f)) { if (b1 < 0.f || jz < a1) a1 = jz; b1 = 0.f; }
              }
              if (first_pass && v0) {
                unsigned char* ap = argo_t + (r * K + c);
                if (kvec2) {
                  *reinterpret_cast<uchar2*>(ap) = make_uchar2((unsigned char)(a0 - mb), (unsigned char)(a1 - mb));
                } else {
                  ap[0] = (unsigned char)(a0 - mb);
                  if (v1) ap[1] = (unsigned char)(a1 - mb);
                }
              }
            }
            if (PROD == P_RO) put_operand(bufH, r, make_float2(b0, b1));
            else *reinterpret_cast<float2*>(&bufH[r * BS + 2 * lane]) = make_float2(b0, b1);
          }
        } else
        for (int r = warp; r < TR; r += NWARP) {
          float b0 = 0.f, b1 = 0.f;
          if (r < rows) {
            const uint4 rec = sRec[r];
            const int m = rec_m(rec), mb = m * N;
            int a0 = mb, a1 = mb;  // tile rows of the arg-max
            const bool flag = rec_binary(rec);
            if (flag && !rec_fallback(rec)) {
              if (MM) rec_max2_padded<BS>(rec, &bufT[2 * lane], b0, b1, a0, a1);
              else rec_max2<BS>(rec, &bufT[2 * lane], r, b0, b1, a0, a1);
              if (rec_has_zero(rec)) {  // a zero entry A[i,jz] contributes the candidate 0*T = 0
                const int jz = mb + rec_jz(rec);
                if (!(b0 > 0.f)) { if (b0 < 0.f || jz < a0) a0 = jz; b0 = 0.f; }
                if (!(b1 > 0.f)) { if (b1 < 0.f || jz < a1) a1 = jz; b1 = 0.f; }
              }
            } else if (flag) {  // binary row with more than kMaxNbr neighbours: walk the bit mask
              const u64* mk = &sRow[(size_t)r * NW];
              int cnt, ar;
              masked_max_first(mk, NW, &bufT[mb * BS + 2 * lane], BS, b0, ar, cnt);
              a0 = mb + ar;
              masked_max_first(mk, NW, &bufT[mb * BS + 2 * lane + 1], BS, b1, ar, cnt);
              a1 = mb + ar;
              if (cnt < N) {
                const int jz = mb + first_zero_bit(mk, NW, N);
                if (!(b0 > 0.f)) { if (b0 < 0.f || jz < a0) a0 = jz; b0 = 0.f; }
                if (!(b1 > 0.f)) { if (b1 < 0.f || jz < a1) a1 = jz; b1 = 0.f; }
              }
            } else {  // general float adjacency (e.g. edge-attention weights): dense scan, exactly the reference order
              const int i = r - mb;
              const float* arow = a.adj + (long long)(mol0 + m) * a.as0 + (long long)i * a.as1;
              const float a00 = __ldg(&arow[0]);
              b0 = a00 * bufT[mb * BS + 2 * lane];
              b1 = a00 * bufT[mb * BS + 2 * lane + 1];
              for (int j = 1; j < N; ++j) {
                const float aj = __ldg(&arow[(long long)j * a.as2]);
                const float2 t = *reinterpret_cast<const float2*>(&bufT[(mb + j) * BS + 2 * lane]);
                const float c0 = aj * t.x, c1 = aj * t.y;
                if (c0 > b0) { b0 = c0; a0 = mb + j; }
                if (c1 > b1) { b1 = c1; a1 = mb + j; }
              }
            }
            if (first_pass && v0) {
              unsigned char* ap = argo_t + (r * K + c);
              if (kvec2) {
                *reinterpret_cast<uchar2*>(ap) = make_uchar2((unsigned char)(a0 - mb), (unsigned char)(a1 - mb));
              } else {
                ap[0] = (unsigned char)(a0 - mb);
                if (v1) ap[1] = (unsigned char)(a1 - mb);
              }
              if (!MM && PROD == P_RO) stg2(prodo_t + (r * K + c), make_float2(b0, b1), v0, v1, kvec2);
            }
          }
          if (PROD == P_RO) put_operand(bufH, r, make_float2(b0, b1));   // readout: the pooled H is the GEMM operand
          else *reinterpret_cast<float2*>(&bufH[r * BS + 2 * lane]) = make_float2(b0, b1);
        }
        if (MM && PROD == P_MID && warp == 0) *reinterpret_cast<float2*>(&bufH[TR * BS + 2 * lane]) = make_float2(0.f, 0.f);   // dummy row: adds nothing
        if (MM && PROD == P_RO) tc::fence_async_smem();
        __syncthreads();
        if (MM && PROD == P_MID && use_mma && tid == 0) {   // T (buf0) is dead now: fetch the weight pieces over it
          const unsigned char* src = reinterpret_cast<const unsigned char*>(a.wimg) + ((size_t)(k0 / KC) * a.ncb + (cb / TC)) * 2 * kWPiece;
          tc::mbar_expect_tx(&bars[0], 2 * kWPiece);
          tc::bulk_g2s(wHi, src, kWPiece, &bars[0]);
          tc::bulk_g2s(wHi + kWPiece, src + kWPiece, kWPiece, &bars[0]);
        }
      } else if (PROD == P_DZ) {
        float2 g = make_float2(0.f, 0.f), c1 = g, c2 = g, mu = g, is = g;
        if (v0) { g.x = __ldg(&a.vec0[c]); c1.x = __ldg(&a.vec1[c]); c2.x = __ldg(&a.vec2[c]); mu.x = __ldg(&a.vec3[c]); is.x = __ldg(&a.vec4[c]); }
        if (v1) { g.y = __ldg(&a.vec0[c + 1]); c1.y = __ldg(&a.vec1[c + 1]); c2.y = __ldg(&a.vec2[c + 1]); mu.y = __ldg(&a.vec3[c + 1]); is.y = __ldg(&a.vec4[c + 1]); }
        float2 colsum = make_float2(0.f, 0.f);
        if (MM && fast4) {
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          // dZ = g*(dY - c1 - (Z - mean)*invstd*c2) folded per channel into dZ = ca*dY + cb*Z + cc (12 coefficient registers instead of 20)
          float4 ca = z4, cb = z4, cc = z4, cs4 = z4;
          if (cv) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(&a.vec0[c4])), c14 = __ldg(reinterpret_cast<const float4*>(&a.vec1[c4]));
            const float4 c24 = __ldg(reinterpret_cast<const float4*>(&a.vec2[c4])), mu4 = __ldg(reinterpret_cast<const float4*>(&a.vec3[c4]));
            const float4 is4 = __ldg(reinterpret_cast<const float4*>(&a.vec4[c4]));
            ca = g4;
            cb = make_float4(-g4.x * is4.x * c24.x, -g4.y * is4.y * c24.y, -g4.z * is4.z * c24.z, -g4.w * is4.w * c24.w);
            cc = make_float4(-fmaf(cb.x, mu4.x, g4.x * c14.x), -fmaf(cb.y, mu4.y, g4.y * c14.y), -fmaf(cb.z, mu4.z, g4.z * c14.z),
                             -fmaf(cb.w, mu4.w, g4.w * c14.w));
          }
          constexpr int HB = RPW / 4;   // two load batches of HB row pairs: 4*HB float4 registers in flight
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            float4 zv[HB], dv[HB];
#pragma unroll
            for (int u = 0; u < HB; ++u) {
              const int r = warp + (2 * (hb * HB + u) + hh) * NWARP;
              const bool ok = r < rows && cv;
              zv[u] = ok ? __ldg(reinterpret_cast<const float4*>(&in1_t[r * K + c4])) : z4;
              dv[u] = ok ? __ldg(reinterpret_cast<const float4*>(&in0_t[r * K + c4])) : z4;
            }
            if (hb == 0) chunk_sync();
#pragma unroll
            for (int u = 0; u < HB; ++u) {
              const int r = warp + (2 * (hb * HB + u) + hh) * NWARP;
              float4 dz = z4;
              if (r < rows && cv) {
                dz.x = fmaf(ca.x, dv[u].x, fmaf(cb.x, zv[u].x, cc.x));
                dz.y = fmaf(ca.y, dv[u].y, fmaf(cb.y, zv[u].y, cc.y));
                dz.z = fmaf(ca.z, dv[u].z, fmaf(cb.z, zv[u].z, cc.z));
                dz.w = fmaf(ca.w, dv[u].w, fmaf(cb.w, zv[u].w, cc.w));
                cs4.x += dz.x; cs4.y += dz.y; cs4.z += dz.z; cs4.w += dz.w;
              }
              put_operand4(r, dz);
            }
          }
          if (first_pass && a.prod_partials) {   // column sums (db): the two half-warps hold the same channels, then across warps
            cs4.x += __shfl_xor_sync(0xffffffffu, cs4.x, 16); cs4.y += __shfl_xor_sync(0xffffffffu, cs4.y, 16);
            cs4.z += __shfl_xor_sync(0xffffffffu, cs4.z, 16); cs4.w += __shfl_xor_sync(0xffffffffu, cs4.w, 16);
            if (hh == 0) *reinterpret_cast<float4*>(&sWarp[(warp * 16 + hl) * 4]) = cs4;
          }
          tc::fence_async_smem();
          __syncthreads();
          if (first_pass && a.prod_partials && warp == 0 && hh == 0 && cv) {
            float4 t = z4;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) { const float4 u = *reinterpret_cast<const float4*>(&sWarp[(w * 16 + hl) * 4]); t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
            *reinterpret_cast<float4*>(&a.prod_partials[(size_t)tile * K + c4]) = t;
          }
        } else {
        constexpr int PB = MM ? 4 : 8;   // rows in flight per lane: 2*PB float2 registers (8 rows spill at 64 registers per thread)
#pragma unroll
        for (int h = 0; h < RPW / PB; ++h) {
          float2 zv[PB], dv[PB];
#pragma unroll
          for (int q = 0; q < PB; ++q) {
            const int r = warp + (h * PB + q) * NWARP;
            const bool ok = r < rows && v0;
            zv[q] = ok ? ldg2(&in1_t[r * K + c], v0, v1, kvec2) : make_float2(0.f, 0.f);
            dv[q] = ok ? ldg2(&in0_t[r * K + c], v0, v1, kvec2) : make_float2(0.f, 0.f);
          }
          if (h == 0) chunk_sync();
#pragma unroll
          for (int q = 0; q < PB; ++q) {
            const int r = warp + (h * PB + q) * NWARP;
            float2 dz = make_float2(0.f, 0.f);
            if (r < rows && v0) {
              dz.x = g.x * (dv[q].x - c1.x - (zv[q].x - mu.x) * is.x * c2.x);
              if (v1) dz.y = g.y * (dv[q].y - c1.y - (zv[q].y - mu.y) * is.y * c2.y);
              if (first_pass) {
                if (!MM) stg2(prodo_t + (r * K + c), dz, v0, v1, kvec2);
                colsum.x += dz.x; colsum.y += dz.y;
              }
            }
            put_operand(buf0, r, dz);
          }
        }
        if (first_pass && a.prod_partials) *reinterpret_cast<float2*>(&sWarp[(warp * 32 + lane) * 2]) = colsum;
        if (MM) tc::fence_async_smem();
        __syncthreads();
        if (first_pass && a.prod_partials && warp == 0 && v0) {
          float2 s = make_float2(0.f, 0.f);
#pragma unroll
          for (int w = 0; w < NWARP; ++w) { const float2 t = *reinterpret_cast<const float2*>(&sWarp[(w * 32 + lane) * 2]); s.x += t.x; s.y += t.y; }
          a.prod_partials[(size_t)tile * K + c] = s.x;
          if (v1) a.prod_partials[(size_t)tile * K + c + 1] = s.y;
        }
        }   // !fast4
      } else if (PROD == P_DU) {
        float2 colsum = make_float2(0.f, 0.f);
        if (MM && fast4) {
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 cs4 = z4;
          float4 dv[RPW / 2];
#pragma unroll
          for (int sI = 0; sI < RPW / 2; ++sI) {
            const int r = warp + (2 * sI + hh) * NWARP;
            dv[sI] = (r < rows && cv) ? __ldg(reinterpret_cast<const float4*>(&in0_t[r * K + c4])) : z4;
          }
          chunk_sync();
#pragma unroll
          for (int sI = 0; sI < RPW / 2; ++sI) {
            const int r = warp + (2 * sI + hh) * NWARP;
            float4 du = z4;
            if (r < rows && cv) {
              const int ob = rec_m(sRec[r]) * K + c4;   // (molecule-of-the-tile, channel) in the per-molecule O / dO tensors
              const float4 ov = __ldg(reinterpret_cast<const float4*>(&o_t[ob])), go = __ldg(reinterpret_cast<const float4*>(&do_t[ob]));
              du.x = go.x * (1.f - ov.x * ov.x) * (1.f - dv[sI].x * dv[sI].x);
              du.y = go.y * (1.f - ov.y * ov.y) * (1.f - dv[sI].y * dv[sI].y);
              du.z = go.z * (1.f - ov.z * ov.z) * (1.f - dv[sI].z * dv[sI].z);
              du.w = go.w * (1.f - ov.w * ov.w) * (1.f - dv[sI].w * dv[sI].w);
              cs4.x += du.x; cs4.y += du.y; cs4.z += du.z; cs4.w += du.w;
            }
            put_operand4(r, du);
          }
          if (first_pass && a.prod_partials) {
            cs4.x += __shfl_xor_sync(0xffffffffu, cs4.x, 16); cs4.y += __shfl_xor_sync(0xffffffffu, cs4.y, 16);
            cs4.z += __shfl_xor_sync(0xffffffffu, cs4.z, 16); cs4.w += __shfl_xor_sync(0xffffffffu, cs4.w, 16);
            if (hh == 0) *reinterpret_cast<float4*>(&sWarp[(warp * 16 + hl) * 4]) = cs4;
          }
          tc::fence_async_smem();
          __syncthreads();
          if (first_pass && a.prod_partials && warp == 0 && hh == 0 && cv) {
            float4 t = z4;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) { const float4 u = *reinterpret_cast<const float4*>(&sWarp[(w * 16 + hl) * 4]); t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
            *reinterpret_cast<float4*>(&a.prod_partials[(size_t)tile * K + c4]) = t;
          }
        } else {
#pragma unroll
        for (int h = 0; h < RPW / 8; ++h) {
          float2 dv[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int r = warp + (h * 8 + q) * NWARP;
            dv[q] = (r < rows && v0) ? ldg2(&in0_t[r * K + c], v0, v1, kvec2) : make_float2(0.f, 0.f);
          }
          if (h == 0) chunk_sync();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int r = warp + (h * 8 + q) * NWARP;
            float2 du = make_float2(0.f, 0.f);
            if (r < rows && v0) {
              const int m = rec_m(sRec[r]);
              const int ob = m * K + c;   // (molecule-of-the-tile, channel) in the per-molecule O / dO tensors
              const float2 ov = ldg2(&o_t[ob], v0, v1, kvec2), go = ldg2(&do_t[ob], v0, v1, kvec2);
              du.x = go.x * (1.f - ov.x * ov.x) * (1.f - dv[q].x * dv[q].x);
              if (v1) du.y = go.y * (1.f - ov.y * ov.y) * (1.f - dv[q].y * dv[q].y);
              if (first_pass) {
                if (!MM) stg2(prodo_t + (r * K + c), du, v0, v1, kvec2);
                colsum.x += du.x; colsum.y += du.y;
              }
            }
            put_operand(buf0, r, du);
          }
        }
        if (first_pass && a.prod_partials) *reinterpret_cast<float2*>(&sWarp[(warp * 32 + lane) * 2]) = colsum;
        if (MM) tc::fence_async_smem();
        __syncthreads();
        if (first_pass && a.prod_partials && warp == 0 && v0) {
          float2 s = make_float2(0.f, 0.f);
#pragma unroll
          for (int w = 0; w < NWARP; ++w) { const float2 t = *reinterpret_cast<const float2*>(&sWarp[(w * 32 + lane) * 2]); s.x += t.x; s.y += t.y; }
          a.prod_partials[(size_t)tile * K + c] = s.x;
          if (v1) a.prod_partials[(size_t)tile * K + c + 1] = s.y;
        }
        }   // !fast4
      }

      if (PROD == P_X || PROD == P_MID) {
        // S = A.H for this chunk (gcn.py:34): over the dead T buffer (FFMA) or into the operand images (MMA)
        if (MM && fast4) {
#pragma unroll
          for (int sI = 0; sI < RPW / 2; ++sI) {
            const int r = warp + (2 * sI + hh) * NWARP;
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows) s4 = rec_sum4_padded<BS>(sRec[r], &buf1[4 * hl]);
            put_operand4(r, s4);
          }
        } else if (tile_simple) {
#pragma unroll 4
          for (int q = 0; q < RPW; ++q) {
            const int r = warp + q * NWARP;
            float2 s = make_float2(0.f, 0.f);
            if (r < rows) s = rec_sum2_padded<BS>(sRec[r], &buf1[2 * lane]);
            put_operand(buf0, r, s);
          }
        } else
        for (int r = warp; r < TR; r += NWARP) {
          float2 s = make_float2(0.f, 0.f);
          if (r < rows) {
            const uint4 rec = sRec[r];
            const int m = rec_m(rec), mb = m * N;
            const bool flag = rec_binary(rec);
            if (flag && !rec_fallback(rec)) {
              s = MM ? rec_sum2_padded<BS>(rec, &buf1[2 * lane]) : rec_sum2<BS>(rec, &buf1[2 * lane], r);
            } else if (flag) {
              s.x = masked_sum(&sRow[(size_t)r * NW], NW, &buf1[mb * BS + 2 * lane], BS);
              s.y = masked_sum(&sRow[(size_t)r * NW], NW, &buf1[mb * BS + 2 * lane + 1], BS);
            } else {
              const float* arow = a.adj + (long long)(mol0 + m) * a.as0 + (long long)(r - mb) * a.as1;
              for (int j = 0; j < N; ++j) {
                const float aj = __ldg(&arow[(long long)j * a.as2]);
                const float2 hv = *reinterpret_cast<const float2*>(&buf1[(mb + j) * BS + 2 * lane]);
                s.x = fmaf(aj, hv.x, s.x);
                s.y = fmaf(aj, hv.y, s.y);
              }
            }
            if (!MM && first_pass && v0 && a.prod_out) stg2(prodo_t + (r * K + c), s, v0, v1, kvec2);
          }
          put_operand(buf0, r, s);
        }
        if (MM) tc::fence_async_smem();
        __syncthreads();
      }
      if (!MM && PROD == P_RO) opnd = buf1;

      // =========================== GEMM chunk ===========================
      if (MM) {
        if (tid == 0) {
          if (use_mma) {
            const int kleft = K - k0;
            const int nks = (kleft >= KC) ? KC / 16 : (kleft + 15) / 16;
            const bool two = a.ncb > 1;
            const uint32_t aHi = tc::smem_u32(imgHi), aLo = tc::smem_u32(imgLo);
            if (!a.skip_gemm) {
              if (two) {   // second column block's weight pieces -> buf1 (dead since the barrier above), overlapping the first block's MMAs
                const unsigned char* src1 = reinterpret_cast<const unsigned char*>(a.wimg) + ((size_t)(k0 / KC) * a.ncb + 1) * 2 * kWPiece;
                tc::mbar_expect_tx(&bars[2], 2 * kWPiece);
                tc::bulk_g2s(buf1, src1, kWPiece, &bars[2]);
                tc::bulk_g2s(reinterpret_cast<unsigned char*>(buf1) + kWPiece, src1 + kWPiece, kWPiece, &bars[2]);
              }
              constexpr uint32_t idesc = tc::make_idesc(tc::FMT_BF16, 128, 128, false, false);
              for (int cbi = 0; cbi < (two ? 2 : 1); ++cbi) {
                if (cbi == 0) tc::mbar_wait(&bars[0], ph_w);
                else tc::mbar_wait(&bars[2], ph_w1);
                tc::fence_after_sync();
                const uint32_t bHi = cbi ? tc::smem_u32(buf1) : tc::smem_u32(wHi), bLo = bHi + kWPiece;
#pragma unroll
                for (int p = 0; p < NPASS; ++p) {
                  const uint32_t ab = (p & 1) ? aLo : aHi, bb = (p & 2) ? bLo : bHi;
                  // (descriptor of K step ks = base + ks * step: only the 14-bit start-address field moves, it cannot carry out)
                  uint64_t da = tc::make_desc(ab, kImgPitch, 128), db = tc::make_desc(bb, 128 * 16, 128);
                  for (int ks = 0; ks < nks; ++ks) {
                    tc::mma_ss<false>(tmem + (uint32_t)(cbi * kTmemCols), da, db, idesc, (k0 > 0 || p > 0 || ks > 0) ? 1u : 0u);
                    da += (uint64_t)((2 * kImgPitch) >> 4);
                    db += (uint64_t)((2 * 128 * 16) >> 4);
                  }
                }
              }
            }
            if (do_dw) {
              // dW[m][n] += sum over the tile's 128 rows of saved[row][m] * produced[row][k0 + n]: both operands MN-major (rows = K),
              // A = the saved image in buf1 (chunk pitch 2048), B = this chunk's hi image (chunk pitch kImgPitch); single bf16 pass
              if (k0 == 0) tc::mbar_wait(&bars[3], ph_s);
              tc::fence_after_sync();
              constexpr uint32_t idw = tc::make_idesc(tc::FMT_BF16, 128, KC, true, true);
              const uint32_t sa = tc::smem_u32(buf1);
              for (int ks = 0; ks < 8; ++ks) {
                const uint64_t da = tc::make_desc(sa + ks * 256, 128, 2048);
                const uint64_t db = tc::make_desc(aHi + ks * 256, 128, kImgPitch);
                tc::mma_ss<false>(tmem + (uint32_t)(kTmemCols + k0), da, db, idw, (tile_iter > 0 || ks > 0) ? 1u : 0u);
              }
            }
            tc::mma_commit(&bars[1]);
          }
          if (use_tma && k0 + KC < K) {   // next channel chunk of Zprev -> buf1: its last readers finished before the barrier above
            tc::fence_async_smem();
            tc::mbar_expect_tx(&bars[5], kZBoxBytes);
            tc::tma_load_2d(buf1, &tm_in0, k0 + KC, (int)row0, &bars[5]);
          }
          if (first_pass && a.prod_img) {
            // the hi image of this chunk is the formatted operand of the weight-gradient GEMM: tile image [K/8][128][8] bf16
            unsigned char* dst = reinterpret_cast<unsigned char*>(a.prod_img) + ((size_t)tile * a.img_chunks + (k0 / 8)) * 2048;
            const int nch = min(8, a.img_chunks - k0 / 8);
            const unsigned char* simg = (a.skip_gemm && !do_dw && ((k0 / KC) & 1)) ? imgLo : imgHi;
            for (int q = 0; q < nch; ++q) tc::bulk_s2g(dst + (size_t)q * 2048, simg + q * kImgPitch, 2048);
            tc::bulk_commit();
          }
        }
        if (use_mma) {
          mma_pending = true;
          if (!a.skip_gemm) { ph_w ^= 1; if (a.ncb > 1) ph_w1 ^= 1; }
          if (do_dw && k0 == 0) ph_s ^= 1;
        }
        if (first_pass && a.prod_img) store_pending = true;
      } else if (!a.skip_gemm) {
#pragma unroll 2
        for (int kk = 0; kk < KC; kk += 4) {
          float4 av[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) av[r] = *reinterpret_cast<const float4*>(&opnd[(r * TYN + ty) * BS + kk]);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 b0 = *reinterpret_cast<const float4*>(&sW[(kk + k4) * TC + tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sW[(kk + k4) * TC + TC / 2 + tx * 4]);
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int r = 0; r < (MM ? 1 : 8); ++r) {
              const float ar = (k4 == 0) ? av[r].x : (k4 == 1) ? av[r].y : (k4 == 2) ? av[r].z : av[r].w;
#pragma unroll
              for (int cc = 0; cc < (MM ? 1 : 8); ++cc) acc[r][cc] = fmaf(ar, bv[cc], acc[r][cc]);
            }
          }
        }
      }
      if (!MM) __syncthreads();  // chunk buffers and sW are free again (MM: the mbarrier / bulk-group waits at the next chunk start cover it)
    }
    if (MM && tid == 0 && store_pending) tc::bulk_wait_read0();
    if (MM) store_pending = false;
    if (a.skip_gemm) break;

    // =========================== EPILOGUE ===========================
    if (MM) {
      // accumulator tile: TMEM -> registers (thread = row) -> bias / activation -> staging tile in shared memory
      if (mma_pending) {
        if (warp == 0) tc::mbar_wait(&bars[1], ph_mma);
        ph_mma ^= 1;
        mma_pending = false;
        // every MMA of this tile has read its operand images (and the bulk store of the hi image was drained above): the image
        // region is free. Persistent forward CTAs request the NEXT tile's first Zprev chunk into it now; it lands under the epilogue.
        if (use_tma && tid == 0 && tix + (int)gridDim.x < a.ntiles) {
          const int ntile = a.reverse ? a.ntiles - 1 - (tix + (int)gridDim.x) : tix + (int)gridDim.x;
          tc::fence_async_smem();
          tc::mbar_expect_tx(&bars[5], kZBoxBytes);
          tc::tma_load_2d(zbox0, &tm_in0, 0, (int)((size_t)ntile * a.MT * N), &bars[5]);
        }
        __syncthreads();
      }
      tc::fence_after_sync();
      constexpr int CPG = TC / (NWARP / 4);   // columns per warp group (the four warps of a group cover the 128 TMEM lanes)
      const int q = warp & 3, hsel = warp >> 2;
      const int row = q * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < CPG; c0 += 32) {
        const int cl = hsel * CPG + c0;
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((cb / TC) * kTmemCols + cl), v);
        if (EPI == E_READOUT) {   // (E_ZSTATS adds its bias in the store pass, where a thread owns four fixed columns)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int gc = cb + cl + i;
            const float bv = (a.bias && gc < NC) ? __ldg(&a.bias[gc]) : 0.f;
            v[i] = (row < rows && gc < NC) ? fast_tanh(v[i] + bv) : 0.f;   // gcn_encoder.py:51
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(&sC[row * TCP + cl + i]) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      if ((EPI == E_DY || EPI == E_ATG_DY) && warp == 0)   // all-zero dummy row of the padded column records
        *reinterpret_cast<float4*>(&sC[TR * TCP + lane * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
      tc::fence_before_sync();
      __syncthreads();
      if (EPI == E_ZSTATS || EPI == E_STORE || EPI == E_READOUT) {
        // coalesced store of the finished tile
        const bool vec_ok = (NC % 4 == 0);
        static_assert(NT % (TC / 4) == 0, "a thread keeps its four columns across the store pass");
        float bb[4] = {0.f, 0.f, 0.f, 0.f};
        if (EPI == E_ZSTATS && a.bias) {
          const int gc0 = cb + (tid % (TC / 4)) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (gc0 + i < NC) bb[i] = __ldg(&a.bias[gc0 + i]);
        }
        for (int idx = tid; idx < rows * (TC / 4); idx += NT) {
          const int r = idx / (TC / 4), cl = (idx - r * (TC / 4)) * 4;
          const int gc = cb + cl;
          if (gc >= NC) continue;
          float4 v = *reinterpret_cast<const float4*>(&sC[r * TCP + cl]);
          v.x += bb[0]; v.y += bb[1]; v.z += bb[2]; v.w += bb[3];
          float* dst = out0_t + (r * NC + gc);
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
      if (EPI == E_ZSTATS) {
        // per-tile (mean, M2) per channel from the staged accumulators (bias-free: the bias shifts the mean only).
        // thread = (row slice, column): ONE pass with the slice's first row as pivot, then a fixed-order Chan merge.
        constexpr int PARTS = NT / TC;   // row slices (2 or 4)
        const int cl = tid & (TC - 1), part = tid / TC;
        const int r_lo = part * (TR / PARTS), r_hi = min(rows, r_lo + TR / PARTS);
        float s = 0.f, q2 = 0.f, pv = 0.f;
        if (r_lo < r_hi) {
          pv = sC[r_lo * TCP + cl];
          for (int r = r_lo + 1; r < r_hi; ++r) { const float d = sC[r * TCP + cl] - pv; s += d; q2 = fmaf(d, d, q2); }
        }
        const float np = (float)max(r_hi - r_lo, 0);
        const float mp = (np > 0.f) ? pv + s / np : 0.f;              // slice mean
        const float m2p = (np > 0.f) ? fmaxf(q2 - s * s / np, 0.f) : 0.f;   // slice M2
        sWarp[part * TC + cl] = mp;
        sWarp[(PARTS + part) * TC + cl] = m2p;
        __syncthreads();
        if (tid < TC && cb + tid < NC) {
          float cn = 0.f, cm = 0.f, cm2 = 0.f;
#pragma unroll
          for (int pp = 0; pp < PARTS; ++pp) {
            const float nt = (float)max(min(rows, (pp + 1) * (TR / PARTS)) - pp * (TR / PARTS), 0);
            if (nt > 0.f) {
              const float totn = cn + nt, dlt = sWarp[pp * TC + tid] - cm;
              cm += dlt * (nt / totn);
              cm2 += sWarp[(PARTS + pp) * TC + tid] + dlt * dlt * (cn * nt / totn);
              cn = totn;
            }
          }
          const float bias_c = a.bias ? __ldg(&a.bias[cb + tid]) : 0.f;
          a.partials[((size_t)tile * 2 + 0) * NC + cb + tid] = cm + bias_c;
          a.partials[((size_t)tile * 2 + 1) * NC + cb + tid] = cm2;
        }
      }
    } else if (EPI == E_ZSTATS || EPI == E_STORE) {
      // thread's output coordinates: rows r*TYN+ty, local cols {tx*4..+3, TC/2+tx*4..+3}
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
        for (int r = 0; r < (MM ? 1 : 8); ++r) {
          const int row = r * TYN + ty;
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[r][(MM ? 0 : h * 4 + q)] += bb[q];
          if (row < rows) {
            float* dst = out0_t + (row * NC + gc);
            if (vec_ok && gc + 3 < NC) {
              *reinterpret_cast<float4*>(dst) = make_float4(acc[r][(MM ? 0 : h * 4 + 0)], acc[r][(MM ? 0 : h * 4 + 1)], acc[r][(MM ? 0 : h * 4 + 2)], acc[r][(MM ? 0 : h * 4 + 3)]);
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (gc + q < NC) dst[q] = acc[r][(MM ? 0 : h * 4 + q)];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) colv[h * 4 + q] += acc[r][(MM ? 0 : h * 4 + q)];
          }
        }
      }
      if (EPI == E_ZSTATS) {
        // per-tile Welford partials (mean, M2) per channel: robust against |mean| >> std, merged in fp64 later
#pragma unroll
        for (int h = 0; h < 2; ++h)
          *reinterpret_cast<float4*>(&sRed[ty * TC + h * (TC / 2) + tx * 4]) = make_float4(colv[h * 4], colv[h * 4 + 1], colv[h * 4 + 2], colv[h * 4 + 3]);
        __syncthreads();
        if (tid < TC) {
          float s = 0.f;
#pragma unroll 8
          for (int y = 0; y < TYN; ++y) s += sRed[y * TC + tid];
          sMean[tid] = s / (float)rows;
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 mu4 = *reinterpret_cast<const float4*>(&sMean[h * (TC / 2) + tx * 4]);
          const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float m2 = 0.f;
#pragma unroll
            for (int r = 0; r < (MM ? 1 : 8); ++r)
              if (r * TYN + ty < rows) { const float d = acc[r][(MM ? 0 : h * 4 + q)] - mu[q]; m2 = fmaf(d, d, m2); }
            colv[h * 4 + q] = m2;
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
          *reinterpret_cast<float4*>(&sRed[ty * TC + h * (TC / 2) + tx * 4]) = make_float4(colv[h * 4], colv[h * 4 + 1], colv[h * 4 + 2], colv[h * 4 + 3]);
        __syncthreads();
        if (tid < TC && cb + tid < NC) {
          float s = 0.f;
#pragma unroll 8
          for (int y = 0; y < TYN; ++y) s += sRed[y * TC + tid];
          a.partials[((size_t)tile * 2 + 0) * NC + cb + tid] = sMean[tid];
          a.partials[((size_t)tile * 2 + 1) * NC + cb + tid] = s;
        }
        __syncthreads();
      }
    } else {  // FFMA, E_READOUT / E_ATG: stage the finished tile in shared memory
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
        for (int r = 0; r < (MM ? 1 : 8); ++r) {
          const int row = r * TYN + ty;
          float v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            v[q] = acc[r][(MM ? 0 : h * 4 + q)];
            if (EPI == E_READOUT) v[q] = (row < rows && gc + q < NC) ? fast_tanh(v[q] + bb[q]) : 0.f;   // gcn_encoder.py:51
          }
          if (EPI == E_READOUT && row < rows) {
            float* dst = out0_t + (row * NC + gc);
            if (NC % 4 == 0 && gc + 3 < NC) {
              *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (gc + q < NC) dst[q] = v[q];
            }
          }
          *reinterpret_cast<float4*>(&sC[row * TCP + cl]) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
      __syncthreads();
    }

    // ---- epilogue work on the staged tile (both engines) ----
    if (EPI == E_READOUT) {
      // O[b,:] = tanh(sum over ALL N rows of D[b,i,:])  (unmasked, gcn_encoder.py:52-53)
      for (int idx = tid; idx < nm * TC; idx += NT) {
        const int m = idx / TC, cl = idx - m * TC;
        if (cb + cl < NC) {
          float s = 0.f;
          for (int i = 0; i < N; ++i) s += sC[(m * N + i) * TCP + cl];
          a.out1[(size_t)(mol0 + m) * NC + cb + cl] = fast_tanh(s);
        }
      }
      __syncthreads();
    } else if (EPI == E_ATG) {
      // dHprev[j,:] = sum_i A[i,j] * dS[i,:]   (column records; 4 channels per lane)
      for (int r = warp; r < rows; r += NWARP) {
        const uint4 crec = sCRec[r];
        const int m = rec_m(crec), mb = m * N;
        const bool flag = rec_binary(crec);
        for (int cl = lane * 4; cl < TC; cl += 128) {
          if (cb + cl >= NC) break;
          float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
          if (flag && !rec_fallback(crec)) {
            s = rec_sum4<TCP>(crec, &sC[cl]);
          } else if (flag) {
            const u64* mk = &sCol[(size_t)r * NW];
            s.x = masked_sum(mk, NW, &sC[mb * TCP + cl], TCP);
            s.y = masked_sum(mk, NW, &sC[mb * TCP + cl + 1], TCP);
            s.z = masked_sum(mk, NW, &sC[mb * TCP + cl + 2], TCP);
            s.w = masked_sum(mk, NW, &sC[mb * TCP + cl + 3], TCP);
          } else {
            const float* acol = a.adj + (long long)(mol0 + m) * a.as0 + (long long)(r - mb) * a.as2;
            for (int i = 0; i < N; ++i) {
              const float ai = __ldg(&acol[(long long)i * a.as1]);
              const float4 dv = *reinterpret_cast<const float4*>(&sC[(mb + i) * TCP + cl]);
              s.x = fmaf(ai, dv.x, s.x); s.y = fmaf(ai, dv.y, s.y); s.z = fmaf(ai, dv.z, s.z); s.w = fmaf(ai, dv.w, s.w);
            }
          }
          float* dst = out0_t + (r * NC + cb + cl);
          if (NC % 4 == 0 && cb + cl + 3 < NC) {
            *reinterpret_cast<float4*>(dst) = s;
          } else {
            const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (cb + cl + q < NC) dst[q] = sv[q];
          }
        }
      }
      __syncthreads();
    } else if (EPI == E_DY || EPI == E_ATG_DY) {
      // ---- fused first half of the previous layer's backward (gcn_encoder.py:43-50 differentiated) ----
      // staged tile -> dH (A^T gather, in place) -> dT = pool scatter through the saved arg-max -> dY = dT*(1-T^2)
      // MM: row TR of the staged tile is the all-zero dummy row of the padded column records, row TR of the arg-max tile holds 0xff
      // (no atom index) so a dummy neighbour never matches
      unsigned char* sArg = reinterpret_cast<unsigned char*>(sC + (MM ? TR + 1 : TR) * TCP);   // [TR (+1)][TC] arg-max of the previous layer
      if (MM && warp == 0) *reinterpret_cast<unsigned*>(&sArg[TR * TC + lane * 4]) = 0xffffffffu;   // (the zero row of sC is written with the staging)
      {
        // TR * (TC / 16) / NT sixteen-byte pieces per thread: all loads first, then the stores
        constexpr int PPT = TR * (TC / 16) / NT;
        static_assert(PPT * NT == TR * (TC / 16), "arg-max staging: whole pieces per thread");
        const bool vec = (NC % 16 == 0);
        uint4 pv[PPT];
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
          const int idx = tid + u * NT;
          const int r = idx / (TC / 16), c16 = (idx - r * (TC / 16)) * 16;
          const int gc = cb + c16;
          pv[u] = make_uint4(0u, 0u, 0u, 0u);
          if (r < rows && gc < NC) {
            const unsigned char* src = argi_t + (r * NC + gc);
            if (vec) {
              pv[u] = __ldg(reinterpret_cast<const uint4*>(src));
            } else {
              unsigned w[4] = {0u, 0u, 0u, 0u};
              for (int b = 0; b < 16; ++b)
                if (gc + b < NC) w[b >> 2] |= (unsigned)__ldg(&src[b]) << ((b & 3) * 8);
              pv[u] = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
          const int idx = tid + u * NT;
          const int r = idx / (TC / 16), c16 = (idx - r * (TC / 16)) * 16;
          *reinterpret_cast<uint4*>(&sArg[r * TC + c16]) = pv[u];
        }
      }
      if (EPI == E_ATG_DY) {
        // dH[j,:] = sum_i A[i,j] * dS[i,:]: 64 columns at a time through registers, written back over dS
#pragma unroll 1
        for (int cs = 0; cs < TC; cs += 64) {
          float2 hv[RPW];
          if (MM && tile_simple && NC % 4 == 0) {   // half-warp per row, four channels per lane
            float4 h4[RPW / 2];
#pragma unroll
            for (int sI = 0; sI < RPW / 2; ++sI) {
              const int r = warp + (2 * sI + hh) * NWARP;
              h4[sI] = (r < rows) ? rec_sum4_padded<TCP>(sCRec[r], &sC[cs + 4 * hl]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncthreads();
#pragma unroll
            for (int sI = 0; sI < RPW / 2; ++sI) {
              const int r = warp + (2 * sI + hh) * NWARP;
              if (r < rows) *reinterpret_cast<float4*>(&sC[r * TCP + cs + 4 * hl]) = h4[sI];
            }
            continue;
          }
          if (tile_simple) {
#pragma unroll
            for (int q = 0; q < RPW; ++q) {
              const int r = warp + q * NWARP;
              hv[q] = (r < rows) ? rec_sum2_padded<TCP>(sCRec[r], &sC[cs + 2 * lane]) : make_float2(0.f, 0.f);
            }
          } else
#pragma unroll
          for (int q = 0; q < RPW; ++q) {
            const int r = warp + q * NWARP;
            float2 sv = make_float2(0.f, 0.f);
            if (r < rows) {
              const uint4 crec = sCRec[r];
              const int m = rec_m(crec), mb = m * N;
              if (rec_binary(crec) && !rec_fallback(crec)) {
                sv = MM ? rec_sum2_padded<TCP>(crec, &sC[cs + 2 * lane]) : rec_sum2<TCP>(crec, &sC[cs + 2 * lane], r);
              } else if (rec_binary(crec)) {
                const u64* mk = &sCol[(size_t)r * NW];
                sv.x = masked_sum(mk, NW, &sC[mb * TCP + cs + 2 * lane], TCP);
                sv.y = masked_sum(mk, NW, &sC[mb * TCP + cs + 2 * lane + 1], TCP);
              } else {
                const float* acol = a.adj + (long long)(mol0 + m) * a.as0 + (long long)(r - mb) * a.as2;
                for (int i = 0; i < N; ++i) {
                  const float ai = __ldg(&acol[(long long)i * a.as1]);
                  const float2 dv = *reinterpret_cast<const float2*>(&sC[(mb + i) * TCP + cs + 2 * lane]);
                  sv.x = fmaf(ai, dv.x, sv.x);
                  sv.y = fmaf(ai, dv.y, sv.y);
                }
              }
            }
            hv[q] = sv;
          }
          __syncthreads();
#pragma unroll
          for (int q = 0; q < RPW; ++q) {
            const int r = warp + q * NWARP;
            if (r < rows) *reinterpret_cast<float2*>(&sC[r * TCP + cs + 2 * lane]) = hv[q];
          }
        }
      }
      __syncthreads();
      {
        const int cl = lane * 4, gc = cb + cl;
        const bool act = cl < TC && gc < NC;
        const bool vec4 = (NC % 4 == 0);
        float scs[4], shs[4], mu[4], is[4], s1[4], s2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool ok = act && gc + e < NC;
          scs[e] = ok ? __ldg(&a.pv_scale[gc + e]) * kTanhScale : 0.f;   // (tanh scale folded in, as in the forward)
          shs[e] = ok ? __ldg(&a.pv_shift[gc + e]) * kTanhScale : 0.f;
          mu[e] = ok ? __ldg(&a.pv_mean[gc + e]) : 0.f;
          is[e] = ok ? __ldg(&a.pv_invstd[gc + e]) : 0.f;
          s1[e] = 0.f;
          s2[e] = 0.f;
        }
        auto load_z = [&](int r) -> float4 {
          float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          if (act && r < rows) {
            const float* zp = zprev_t + (r * NC + gc);
            if (vec4) z = __ldg(reinterpret_cast<const float4*>(zp));
            else { z.x = __ldg(zp); if (gc + 1 < NC) z.y = __ldg(zp + 1); if (gc + 2 < NC) z.z = __ldg(zp + 2); if (gc + 3 < NC) z.w = __ldg(zp + 3); }
          }
          return z;
        };
        auto dy_row = [&](int r, const float4 z4) {
          const uint4 crec = sCRec[r];
          const int m = rec_m(crec), mb = m * N;
          const unsigned jrel = (unsigned)(r - mb);
          float dt[4] = {0.f, 0.f, 0.f, 0.f};
          if (rec_binary(crec) && !rec_fallback(crec)) {
            // dT[j,c] = sum over rows i that list j as a neighbour and whose arg-max in channel c is j
            const int cnt = rec_cnt(crec);
            // one XOR per neighbour turns "arg-max byte == my index" into "byte of x is zero": a single logic op with predicate
            // output per channel instead of extract + compare
            const unsigned jrel4 = jrel * 0x01010101u;
#define OCGCN_SCAT_ADD(G, AG, OK)                                                             \
            {                                                                                 \
              const unsigned x_ = (AG) ^ jrel4;                                               \
              if (OK && !(x_ & 0x000000ffu)) dt[0] += G.x;                                    \
              if (OK && !(x_ & 0x0000ff00u)) dt[1] += G.y;                                    \
              if (OK && !(x_ & 0x00ff0000u)) dt[2] += G.z;                                    \
              if (OK && !(x_ & 0xff000000u)) dt[3] += G.w;                                    \
            }
            {   // first four neighbours without branches: all eight shared-memory loads in flight at once
              // (MM: padded records — unused slots point at the dummy rows, no count checks; FFMA engine: re-read a valid row, masked)
              const int i0 = (MM || cnt > 0) ? rec_idx<0>(crec) : r, i1 = (MM || cnt > 1) ? rec_idx<1>(crec) : i0;
              const int i2 = (MM || cnt > 2) ? rec_idx<2>(crec) : i0, i3 = (MM || cnt > 3) ? rec_idx<3>(crec) : i0;
              const float4 g0 = *reinterpret_cast<const float4*>(&sC[i0 * TCP + cl]);
              const float4 g1 = *reinterpret_cast<const float4*>(&sC[i1 * TCP + cl]);
              const float4 g2 = *reinterpret_cast<const float4*>(&sC[i2 * TCP + cl]);
              const float4 g3 = *reinterpret_cast<const float4*>(&sC[i3 * TCP + cl]);
              const unsigned q0 = *reinterpret_cast<const unsigned*>(&sArg[i0 * TC + cl]);
              const unsigned q1 = *reinterpret_cast<const unsigned*>(&sArg[i1 * TC + cl]);
              const unsigned q2 = *reinterpret_cast<const unsigned*>(&sArg[i2 * TC + cl]);
              const unsigned q3 = *reinterpret_cast<const unsigned*>(&sArg[i3 * TC + cl]);
              const bool k0 = MM || cnt > 0, k1 = MM || cnt > 1, k2 = MM || cnt > 2, k3 = MM || cnt > 3;
              OCGCN_SCAT_ADD(g0, q0, k0) OCGCN_SCAT_ADD(g1, q1, k1) OCGCN_SCAT_ADD(g2, q2, k2) OCGCN_SCAT_ADD(g3, q3, k3)
            }
#define OCGCN_SCAT_STEP(Q)                                                                    \
            if (cnt > Q) {                                                                    \
              const int ir = rec_idx<Q>(crec);                                                \
              const float4 g = *reinterpret_cast<const float4*>(&sC[ir * TCP + cl]);          \
              const unsigned ag = *reinterpret_cast<const unsigned*>(&sArg[ir * TC + cl]);    \
              OCGCN_SCAT_ADD(g, ag, true)                                                     \
            }
            if (cnt > 4) {
              OCGCN_SCAT_STEP(4) OCGCN_SCAT_STEP(5) OCGCN_SCAT_STEP(6) OCGCN_SCAT_STEP(7) OCGCN_SCAT_STEP(8)
              if (cnt > 9) { OCGCN_SCAT_STEP(9) OCGCN_SCAT_STEP(10) OCGCN_SCAT_STEP(11) OCGCN_SCAT_STEP(12) }
            }
#undef OCGCN_SCAT_STEP
#undef OCGCN_SCAT_ADD
          } else if (rec_binary(crec)) {
            const u64* mk = &sCol[(size_t)r * NW];
            for (int w = 0; w < NW; ++w) {
              u64 bits = mk[w];
              while (bits) {
                const int ir = mb + 64 * w + __ffsll((long long)bits) - 1;
                bits &= bits - 1;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (sArg[ir * TC + cl + e] == jrel) dt[e] += sC[ir * TCP + cl + e];
              }
            }
          } else {
            const float* acol = a.adj + (long long)(mol0 + m) * a.as0 + (long long)(r - mb) * a.as2;
            for (int i = 0; i < N; ++i) {
              const float ai = __ldg(&acol[(long long)i * a.as1]);
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (sArg[(mb + i) * TC + cl + e] == jrel) dt[e] = fmaf(ai, sC[(mb + i) * TCP + cl + e], dt[e]);
            }
          }
          const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
          float dy[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float t = fast_tanh_scaled(fmaf(zz[e], scs[e], shs[e]));
            dy[e] = (gc + e < NC) ? dt[e] * (1.f - t * t) : 0.f;
            s1[e] += dy[e];
            s2[e] = fmaf(dy[e], (zz[e] - mu[e]) * is[e], s2[e]);
          }
          float* dst = out0_t + (r * NC + gc);
          if (vec4) {
            *reinterpret_cast<float4*>(dst) = make_float4(dy[0], dy[1], dy[2], dy[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (gc + e < NC) dst[e] = dy[e];
          }
        };
        // Z rows are fetched two rows (of this warp) ahead of their use through a rotating register queue. Deeper queues were
        // measured slower at 64 registers per thread (4 deep: spills inside this loop, 2.16 vs 2.11 ms per step at C3)
        float4 zq0 = load_z(warp), zq1 = load_z(warp + NWARP);
#pragma unroll 1
        for (int r = warp; r < rows; r += NWARP) {
          const float4 zc = zq0;
          zq0 = zq1;
          zq1 = load_z(r + 2 * NWARP);
          if (act) dy_row(r, zc);
        }
        __syncthreads();   // every warp is done with the dH tile: reuse it for the cross-warp sums
        float* sRed = sC;    // [NWARP][2][TC]
        if (cl < TC) {
          *reinterpret_cast<float4*>(&sRed[(warp * 2 + 0) * TC + cl]) = make_float4(s1[0], s1[1], s1[2], s1[3]);
          *reinterpret_cast<float4*>(&sRed[(warp * 2 + 1) * TC + cl]) = make_float4(s2[0], s2[1], s2[2], s2[3]);
        }
        __syncthreads();
        if (tid < TC && cb + tid < NC) {
          float t1 = 0.f, t2 = 0.f;
#pragma unroll
          for (int w = 0; w < NWARP; ++w) { t1 += sRed[(w * 2 + 0) * TC + tid]; t2 += sRed[(w * 2 + 1) * TC + tid]; }
          a.partials[((size_t)tile * 2 + 0) * NC + cb + tid] = t1;
          a.partials[((size_t)tile * 2 + 1) * NC + cb + tid] = t2;
        }
        __syncthreads();
      }
    } else if (MM) {
      __syncthreads();   // the staging tile aliases the chunk buffers of the next column pass
    }
  }   // column passes
  __syncthreads();   // the next tile restages the neighbour records and reuses every buffer
  }   // tiles of this CTA
  pdl_trigger();

  if (do_dw) {
    // this CTA's partial weight gradient: TMEM columns [128, 256) -> dw_partial[cta][m][n] (merged in fixed order afterwards)
    if (mma_pending) {
      tc::mbar_wait(&bars[1], ph_mma);
      ph_mma ^= 1;
      mma_pending = false;
    }
    tc::fence_after_sync();
    constexpr int CPG = TC / (NWARP / 4);
    const int q = warp & 3, hsel = warp >> 2;
    const int m = q * 32 + lane;
    float* outp = a.dw_partial + (size_t)blockIdx.x * a.dw_M * K;
    const bool vec4 = (K % 4 == 0);
#pragma unroll 1
    for (int c0 = 0; c0 < CPG; c0 += 32) {
      const int cl = hsel * CPG + c0;
      float v[32];
      tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(kTmemCols + cl), v);
      if (blockIdx.x < (unsigned)a.ntiles && m < a.dw_M && cl < K) {
        float* dst = outp + (size_t)m * K + cl;
        if (vec4 && cl + 32 <= K) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cl + i < K) dst[i] = v[i];
        }
      }
    }
  }
  if (MM && use_mma) {
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, tcols);
  }
}

}  // namespace ocgcn
