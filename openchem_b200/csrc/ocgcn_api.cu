// extern "C" boundary of libocgcn.so: argument validation, workspace layout, kernel sequencing.
// No torch types, no device allocation, no synchronisation: everything is enqueued on the caller's stream.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "ocgcn_aux.cuh"
#include "ocgcn_common.cuh"
#include "ocgcn_compact.cuh"
#include "ocgcn_dw.cuh"
#include "ocgcn_fused.cuh"
#include "ocgcn_head.cuh"
#include "ocgcn_pool.cuh"
#include "ocgcn_tile.cuh"

namespace ocgcn {

static thread_local int tl_tile_parity = 0;
static thread_local int tl_launches = 0;
static thread_local int tl_cuda_err = 0;

// ---- opt-in per-kernel timing (bench.py's roofline leg): CUDA events on the launching stream ----------
enum KernelKind : int {
  KK_ADJ_PACK = 0, KK_TRANSPOSE, KK_FWD_FIRST, KK_FWD_MID, KK_BN_FINALIZE, KK_FWD_READOUT, KK_BWD_READOUT, KK_BWD_DY,
  KK_BWD_FINALIZE, KK_BWD_DZ, KK_GEMM_TN, KK_REDUCE, KK_BN_APPLY, KK_WPREP, KK_DW_MMA, KK_EXPAND, KK_PACK, KK_HEAD_FWD, KK_HEAD_LOSS, KK_HEAD_BWD, KK_POOL_FWD, KK_POOL_BWD, KK_FUSED_EVAL, KK_FUSED_TRAIN, KK_COUNT
};
static const char* const kKindNames[KK_COUNT] = {"adj_pack", "transpose", "fwd_first", "fwd_mid", "bn_finalize", "fwd_readout",
                                                 "bwd_readout", "bwd_dy", "bwd_finalize", "bwd_dz", "gemm_tn", "reduce", "bn_apply",
                                                 "wprep", "dw_mma", "expand_compact", "pack_compact", "head_fwd", "head_loss", "head_bwd", "pool_fwd", "pool_bwd", "fused_eval", "fused_train_fwd"};
struct ProfRec { int kind; cudaEvent_t e0, e1; };
// process-wide (autograd runs backward on its own thread), guarded by a mutex; off by default
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
struct ProfScope {
  cudaStream_t st; cudaEvent_t e1; bool on;
  ProfScope(int kind, cudaStream_t s) : st(s), e1(nullptr), on(false) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec r; r.kind = kind;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, st);
    e1 = r.e1; on = true;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
  }
  ~ProfScope() { if (on) cudaEventRecord(e1, st); }
};

#define CU_TRY(expr)                                   \
  do {                                                 \
    cudaError_t e__ = (expr);                          \
    if (e__ != cudaSuccess) {                          \
      tl_cuda_err = (int)e__;                          \
      return OCGCN_ECUDA;                              \
    }                                                  \
  } while (0)
#define ST_TRY(expr)              \
  do {                            \
    int s__ = (expr);             \
    if (s__ != OCGCN_OK) return s__; \
  } while (0)
#define LAUNCH_CHECK()       \
  do {                       \
    ++tl_launches;           \
    CU_TRY(cudaGetLastError()); \
  } while (0)

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// Launch one kernel of a forward / backward call's chain. With programmatic dependent launch enabled (default; OCGCN_PDL=0
// turns it off) the kernel may start its prologue while its predecessor on the stream drains; every kernel launched through
// here calls pdl_wait() before its first global-memory access (ocgcn_common.cuh). Captured into CUDA graphs as programmatic edges.
static bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("OCGCN_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
static bool tma_enabled() {
  static const bool on = [] { const char* e = getenv("OCGCN_TMA"); return !(e && e[0] == '0'); }();
  return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  if (pdl_enabled()) { cfg.attrs = attr; cfg.numAttrs = 1; }
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
constexpr int kMaxPersistent = 2 * 148;   // tensor-core tile kernels: two resident CTAs per SM loop over the tiles

// ---------------------------------------------------------------------------------------------------------
// workspace layout (byte offsets). "saved" lives from forward to backward, "scratch" only within one call.
// ---------------------------------------------------------------------------------------------------------
struct Layout {
  int TR, MT, NW, tiles, Fmax;
  int mm;                            // 1: tensor-core engine (bf16 operand images), 0: exact-fp32 FFMA engine
  int fuse_dw;                       // mm and every width <= 128: weight gradients are accumulated inside the backward tile kernels
  int KP[OCGCN_MAX_LAYERS + 1], KPE;  // channel counts padded to multiples of 64 (operand image extents)
  size_t rows;
  // saved
  size_t rowmask, colmask, molflag, nbr, cnbr;
  size_t Z[OCGCN_MAX_LAYERS], arg[OCGCN_MAX_LAYERS], S[OCGCN_MAX_LAYERS], stats[OCGCN_MAX_LAYERS];
  size_t HL, D, O;
  size_t saved_total;
  // scratch
  size_t partials, prodpart[OCGCN_MAX_LAYERS + 1], WT[OCGCN_MAX_LAYERS], WdT, G[3], dU, coef, gemm_partial[OCGCN_MAX_LAYERS + 1];   // [L] = readout
  size_t Wf[OCGCN_MAX_LAYERS], Wro, Wb[OCGCN_MAX_LAYERS], Wrob, dZimg, dUimg;   // mm: weight operand images, dZ / dU tile images
  size_t scratch_total;
  int max_splits;
};

static void gemm_tn_config(size_t R, int M, int Nn, int* splits, int* rows_per_split) {
  const int nblocks = ((M + 127) / 128) * ((Nn + 127) / 128);
  int s = 296 / nblocks;
  if (s < 1) s = 1;
  size_t rps = (R + s - 1) / s;
  if (rps < 4 * GK) rps = 4 * GK;
  rps = (rps + GK - 1) / GK * GK;
  *rows_per_split = (int)rps;
  *splits = (int)((R + rps - 1) / rps);
}

static int make_layout(const ocgcn_desc* d, bool layer_only, Layout* lo) {
  memset(lo, 0, sizeof(*lo));
  if (!d) return OCGCN_EINVAL;
  if (d->B <= 0 || d->N <= 0 || d->L <= 0 || d->L > OCGCN_MAX_LAYERS) return OCGCN_EINVAL;
  if (d->N > OCGCN_MAX_ATOMS) return OCGCN_EUNSUPPORTED;
  if (!layer_only && d->E <= 0) return OCGCN_EINVAL;
  if (layer_only && d->L != 1) return OCGCN_EINVAL;
  for (int l = 0; l <= d->L; ++l)
    if (d->F[l] <= 0) return OCGCN_EINVAL;
  if (d->mode != OCGCN_FP32 && d->mode != OCGCN_TF32 && d->mode != OCGCN_BF16) return OCGCN_EINVAL;
  if ((size_t)d->B * d->N >= (size_t)1 << 31) return OCGCN_EUNSUPPORTED;
  lo->TR = d->N <= 128 ? 128 : 256;
  lo->MT = lo->TR / d->N;
  lo->NW = (d->N + 63) / 64;
  lo->tiles = (d->B + lo->MT - 1) / lo->MT;
  lo->rows = (size_t)d->B * d->N;
  int Fmax = d->E > 0 && !layer_only ? d->E : 1;
  for (int l = 0; l <= d->L; ++l) Fmax = d->F[l] > Fmax ? d->F[l] : Fmax;
  lo->Fmax = Fmax;
  // the tensor-core engine covers 128-row tiles and widths up to 256; anything else runs the exact-fp32 engine
  lo->mm = (d->mode != OCGCN_FP32 && d->N <= 128 && Fmax <= 256) ? 1 : 0;
  int KPmax = 64;
  for (int l = 0; l <= d->L; ++l) { lo->KP[l] = (d->F[l] + 63) / 64 * 64; KPmax = lo->KP[l] > KPmax ? lo->KP[l] : KPmax; }
  lo->KPE = layer_only ? 64 : (d->E + 63) / 64 * 64;
  lo->fuse_dw = (lo->mm && Fmax <= 128) ? 1 : 0;
  const size_t rows = lo->rows;
  size_t off = 0;
  auto take = [&off](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
  lo->rowmask = take(rows * lo->NW * 8);
  lo->colmask = take(rows * lo->NW * 8);
  lo->molflag = take((size_t)d->B * 4);
  lo->nbr = take(rows * 16);
  lo->cnbr = take(rows * 16);
  for (int l = 0; l < d->L; ++l) {
    lo->Z[l] = take(rows * d->F[l + 1] * 4);
    lo->arg[l] = layer_only ? 0 : take(rows * d->F[l + 1]);
    lo->S[l] = lo->mm ? take((size_t)lo->tiles * lo->KP[l] * 256) : take(rows * d->F[l] * 4);
    lo->stats[l] = take((size_t)4 * d->F[l + 1] * 4);
  }
  if (lo->mm) {   // backward weight operand images: written by the forward call (same weights), read by the backward call
    for (int l = 0; l < d->L; ++l) lo->Wb[l] = take(wimg_bytes(d->F[l + 1], d->F[l], 2));
    if (!layer_only) lo->Wrob = take(wimg_bytes(d->E, d->F[d->L], 2));
  }
  if (!layer_only) {
    lo->HL = lo->mm ? take((size_t)lo->tiles * lo->KP[d->L] * 256) : take(rows * d->F[d->L] * 4);
    lo->D = take(rows * d->E * 4);
    lo->O = take((size_t)d->B * d->E * 4);
  }
  lo->saved_total = off;

  off = 0;
  lo->partials = take((size_t)lo->tiles * 2 * Fmax * 4);
  for (int l = 0; l <= d->L; ++l) lo->prodpart[l] = take((size_t)lo->tiles * Fmax * 4);   // per layer: merged by ONE reduction launch at the end
  int max_splits = 1;
  size_t max_partial = 0;
  for (int l = 0; l < d->L; ++l) {
    lo->WT[l] = take((size_t)d->F[l] * d->F[l + 1] * 4);
    int s, rps;
    gemm_tn_config(rows, d->F[l], d->F[l + 1], &s, &rps);
    size_t p = (size_t)s * d->F[l] * d->F[l + 1] * 4;
    max_partial = p > max_partial ? p : max_partial;
    max_splits = s > max_splits ? s : max_splits;
  }
  if (!layer_only) {
    lo->WdT = take((size_t)d->F[d->L] * d->E * 4);
    int s, rps;
    gemm_tn_config(rows, d->E, d->F[d->L], &s, &rps);
    size_t p = (size_t)s * d->E * d->F[d->L] * 4;
    max_partial = p > max_partial ? p : max_partial;
    lo->dU = take(rows * d->E * 4);
  }
  for (int i = 0; i < 3; ++i) lo->G[i] = take(rows * Fmax * 4);
  lo->coef = take((size_t)3 * ((Fmax + 3) / 4 * 4) * 4);   // three vectors at a 16-byte aligned stride (float4 loads)
  if (lo->mm) {
    for (int l = 0; l < d->L; ++l) {
      lo->Wf[l] = take(wimg_bytes(d->F[l], d->F[l + 1], 2));
    }
    if (!layer_only) {
      lo->Wro = take(wimg_bytes(d->F[d->L], d->E, 2));
      lo->dUimg = take((size_t)lo->tiles * lo->KPE * 256);
    }
    lo->dZimg = take((size_t)lo->tiles * KPmax * 256);
    const size_t p = (size_t)kMaxPersistent * KPmax * (lo->KPE > KPmax ? lo->KPE : KPmax) * 4;   // per-CTA dW partials (dw_mma_kernel / fused)
    max_partial = p > max_partial ? p : max_partial;
  }
  for (int l = 0; l <= d->L; ++l) lo->gemm_partial[l] = take(max_partial);
  lo->max_splits = max_splits;
  lo->scratch_total = off;
  return OCGCN_OK;
}

constexpr int kMaxDevices = 64;
static int check_device() {
  static std::atomic<int> major_of[kMaxDevices];   // 0 = not asked yet (immutable per device afterwards)
  int dev = 0, major = 0;
  CU_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < kMaxDevices) major = major_of[dev].load(std::memory_order_relaxed);
  if (major == 0) {
    CU_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (dev >= 0 && dev < kMaxDevices) major_of[dev].store(major, std::memory_order_relaxed);
  }
  if (major != 10) return OCGCN_EARCH;
  return OCGCN_OK;
}

// Is `p` device (or managed) memory? cudaPointerGetAttributes costs ~1 us; a training loop passes the same few dozen parameter
// and workspace addresses every step, so validated addresses are remembered in a small per-thread direct-mapped table. (Unified
// virtual addressing: an address that was device memory once can only ever be device memory again, never a host allocation.)
static int check_device_ptr(const void* p) {
  if (!p) return OCGCN_EINVAL;
  if (((uintptr_t)p & 3) != 0) return OCGCN_EALIGN;
  static thread_local const void* seen[256];
  const size_t slot = ((uintptr_t)p >> 4) * 0x9E3779B97F4A7C15ull >> 56;
  if (seen[slot] == p) return OCGCN_OK;
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return OCGCN_EINVAL;
  }
  if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) return OCGCN_EINVAL;  // no CPU fallback
  seen[slot] = p;
  return OCGCN_OK;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) and size class instead of on every launch
template <auto Kern>
static int ensure_dynamic_smem(size_t smem) {
  static std::atomic<size_t> granted[kMaxDevices];
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  const bool track = dev >= 0 && dev < kMaxDevices;
  if (track && granted[dev].load(std::memory_order_relaxed) >= smem) return OCGCN_OK;
  CU_TRY(cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (track) granted[dev].store(smem, std::memory_order_relaxed);
  return OCGCN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------------------
static int mm_grid(int tiles) { return tiles < kMaxPersistent ? tiles : kMaxPersistent; }
// cooperative launch (grid-wide barrier inside the kernel: every CTA must be resident)
template <typename... KArgs, typename... Args>
static cudaError_t launch_cooperative(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
static bool fused_layer_enabled() {
  static const bool on = [] { const char* e = getenv("OCGCN_FUSED_LAYER"); return e && e[0] == '1'; }();   // experimental: off unless asked for
  return on;
}
static bool fused_train_enabled() {
  static const bool on = [] { const char* e = getenv("OCGCN_FUSED_TRAIN"); return !(e && e[0] == '0'); }();
  return on;
}

// ---- TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// row-major fp32 matrix [rows][cols] -> map with a 128-row x 64-column box, no swizzle (dense [128][64] image in shared memory)
static bool make_chunk_map(const float* base, size_t rows, int cols, CUtensorMap* tm) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || rows < 128 || cols < 64 || (cols % 4) != 0 || ((uintptr_t)base & 15)) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {64u, 128u};
  const cuuint32_t estride[2] = {1u, 1u};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int TR, int PROD, int EPI, int MM>
static int launch_tile_t(const TileArgs& a_in, int tiles, cudaStream_t st) {
  TileArgs a = a_in;
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  // forward producers of the tensor-core engine read Zprev chunk by chunk: stage it with TMA when the shape allows (one column block,
  // so buf1 is not needed for a second weight image; at least one full box in each dimension)
  a.use_tma = (MM && (PROD == P_MID || PROD == P_RO) && a.ncb == 1 && tma_enabled() &&
               make_chunk_map(a.in0, (size_t)a.B * a.N, a.K, &tm)) ? 1 : 0;
  // consecutive tile launches of a call alternate the direction they walk the batch in (thread-local parity, reset per API call)
  static const bool zigzag = [] { const char* e = getenv("OCGCN_ZIGZAG"); return !(e && e[0] == '0'); }();
  a.reverse = (MM && zigzag && tiles > kMaxPersistent) ? (tl_tile_parity++ & 1) : 0;   // (batches whose tiles are all co-resident gain nothing)
  constexpr int TC = tile_cols<TR>();
  constexpr int NT = MM ? kThreadsMM : kThreads;
  ProfScope prof(PROD == P_X ? KK_FWD_FIRST : PROD == P_MID ? KK_FWD_MID : PROD == P_RO ? KK_FWD_READOUT : PROD == P_DZ ? KK_BWD_DZ : KK_BWD_READOUT, st);
  auto kern = tile_kernel<TR, TC, PROD, EPI, MM, NT>;
  const size_t smem = tile_smem_bytes<TR, TC, MM, NT>(a.NW);
  ST_TRY((ensure_dynamic_smem<tile_kernel<TR, TC, PROD, EPI, MM, NT>>(smem)));
  // persistent (two CTAs per SM looping over tiles) only where the CTA accumulates the weight gradient across its tiles;
  // everything else launches one CTA per tile (dynamic scheduling balances and de-phases the co-resident CTAs better)
  // forward kernels with the TMA-staged input run persistent too: a CTA requests its next tile's first input chunk during the epilogue
  static const bool persist_fwd = [] { const char* e = getenv("OCGCN_PERSIST_FWD"); return !(e && e[0] == '0'); }();
  CU_TRY(launch_chain(kern, dim3((MM && (a.dw_img || (persist_fwd && a.use_tma))) ? mm_grid(tiles) : tiles), dim3(NT), smem, st, a, tm));
  LAUNCH_CHECK();
  return OCGCN_OK;
}
template <int PROD, int EPI>
static int launch_tile(const Layout& lo, const TileArgs& a, cudaStream_t st) {
  if (lo.mm) return launch_tile_t<128, PROD, EPI, 1>(a, lo.tiles, st);
  if (lo.TR == 128) return launch_tile_t<128, PROD, EPI, 0>(a, lo.tiles, st);
  return launch_tile_t<256, PROD, EPI, 0>(a, lo.tiles, st);
}

// dW (M x Nn) = sum over tiles of Aimg^T . Bimg on tcgen05 (per-CTA partials in `partial`, merged by run_reduce)
static int run_dw_mma(const Layout& lo, const void* Aimg, const void* Bimg, int KPa, int KPb, int M, int Nn, float* partial,
                      cudaStream_t st, int* splits_out) {
  ProfScope prof(KK_DW_MMA, st);
  DwArgs g;
  g.Aimg = Aimg; g.Bimg = Bimg; g.tiles = lo.tiles; g.chA = KPa / 8; g.chB = KPb / 8; g.M = M; g.Nn = Nn; g.partial = partial;
  int splits = lo.tiles < 148 ? lo.tiles : 148;
  g.tiles_per_cta = (lo.tiles + splits - 1) / splits;
  splits = (lo.tiles + g.tiles_per_cta - 1) / g.tiles_per_cta;
  g.rows_per_stage = 128;
  size_t stage = dw_stage_bytes(g.chA, g.chB, 128);
  if ((size_t)(200 * 1024) / stage < 2) {   // wide layers: half-tile stages keep the copy / MMA pipeline at least three deep
    g.rows_per_stage = 64;
    stage = dw_stage_bytes(g.chA, g.chB, 64);
  }
  int stages = (int)((size_t)(200 * 1024) / stage);
  stages = stages < 1 ? 1 : stages > 4 ? 4 : stages;
  g.stages = stages;
  const size_t smem = stage * stages + (size_t)(2 * stages + 1) * 8 + 16;
  ST_TRY((ensure_dynamic_smem<dw_mma_kernel>(smem)));
  CU_TRY(launch_chain(dw_mma_kernel, dim3(splits), dim3(128), smem, st, g));
  LAUNCH_CHECK();
  *splits_out = splits;
  return OCGCN_OK;
}

static int run_wprep(const WPrepJobs& jobs, int n, cudaStream_t st) {
  if (n == 0) return OCGCN_OK;
  ProfScope prof(KK_WPREP, st);
  int pieces = 1;
  for (int j = 0; j < n; ++j) {
    const int p = ((jobs.K[j] + 63) / 64) * ((jobs.Nn[j] + 127) / 128);
    pieces = p > pieces ? p : pieces;
  }
  CU_TRY(launch_chain(wprep_kernel, dim3(8 * pieces, n), dim3(256), 0, st, jobs));
  LAUNCH_CHECK();
  return OCGCN_OK;
}
static void add_wprep(WPrepJobs& jobs, int& n, const float* src, void* dst, int K, int Nn, int form, int parts) {
  jobs.src[n] = src; jobs.dst[n] = dst; jobs.K[n] = K; jobs.Nn[n] = Nn; jobs.form[n] = form; jobs.parts[n] = parts;
  ++n;
}

static int launch_bwd_dy(const Layout& lo, const BwdDyArgs& a, cudaStream_t st) {
  ProfScope prof(KK_BWD_DY, st);
  if (lo.TR == 128) {
    const size_t smem = bwd_dy_smem_bytes<128>(a.NW);
    ST_TRY((ensure_dynamic_smem<bwd_dy_kernel<128>>(smem)));
    bwd_dy_kernel<128><<<lo.tiles, kThreads, smem, st>>>(a);
  } else {
    const size_t smem = bwd_dy_smem_bytes<256>(a.NW);
    ST_TRY((ensure_dynamic_smem<bwd_dy_kernel<256>>(smem)));
    bwd_dy_kernel<256><<<lo.tiles, kThreads, smem, st>>>(a);
  }
  LAUNCH_CHECK();
  return OCGCN_OK;
}

static TileArgs base_tile_args(const ocgcn_desc* d, const Layout& lo, const float* adj, char* saved) {
  TileArgs a;
  memset(&a, 0, sizeof(a));
  a.B = d->B; a.N = d->N; a.NW = lo.NW; a.MT = lo.MT; a.ntiles = lo.tiles;
  a.adj = adj; a.as0 = d->adj_stride[0]; a.as1 = d->adj_stride[1]; a.as2 = d->adj_stride[2];
  a.rowmask = reinterpret_cast<const u64*>(saved + lo.rowmask);
  a.colmask = reinterpret_cast<const u64*>(saved + lo.colmask);
  a.molflag = reinterpret_cast<const unsigned*>(saved + lo.molflag);
  a.nbr = reinterpret_cast<const uint4*>(saved + lo.nbr);
  a.cnbr = reinterpret_cast<const uint4*>(saved + lo.cnbr);
  return a;
}

// The batch as the caller handed it over: dense fp32 x / adj with the strides of the descriptor, or the bit words of a compact batch
// (ocgcn_*_compact entry points; exact 0/1 matrices, ocgcn_pack_compact's layout).
struct GraphIn {
  const float* x;
  const float* adj;
  const uint32_t* x_bits;
  const uint32_t* adj_bits;
  bool compact() const { return adj_bits != nullptr; }
};

static int check_graph_in(const GraphIn& gi) {
  if (gi.compact()) {
    if (!gi.x_bits) return OCGCN_EINVAL;
    ST_TRY(check_device_ptr(gi.adj_bits));
    ST_TRY(check_device_ptr(gi.x_bits));
    if (((uintptr_t)gi.adj_bits & 3) || ((uintptr_t)gi.x_bits & 3)) return OCGCN_EALIGN;
    return OCGCN_OK;
  }
  ST_TRY(check_device_ptr(gi.x));
  ST_TRY(check_device_ptr(gi.adj));
  return OCGCN_OK;
}

static int launch_adj_pack(const ocgcn_desc* d, const GraphIn& gi, int NW, u64* rowmask, u64* colmask, unsigned* molflag, uint4* nbr,
                           uint4* cnbr, cudaStream_t st) {
  ProfScope prof(KK_ADJ_PACK, st);
  const size_t smem = (size_t)2 * d->N * NW * 8;
  const float* no_adj = nullptr;
  const uint32_t* no_bits = nullptr;
  const long long s0 = d->adj_stride[0], s1 = d->adj_stride[1], s2 = d->adj_stride[2];
  if (NW == 1 && d->B >= 1024) {      // N <= 64, large batch: one warp per molecule (small batches need the CTA kernel's parallelism per molecule)
    const dim3 grid((d->B + kWarps - 1) / kWarps);
    if (gi.compact())
      CU_TRY(launch_chain(adj_pack_warp_kernel<true>, grid, dim3(kThreads), 0, st, no_adj, 0ll, 0ll, 0ll, gi.adj_bits, d->B, d->N, rowmask, colmask,
                          molflag, nbr, cnbr));
    else
      CU_TRY(launch_chain(adj_pack_warp_kernel<false>, grid, dim3(kThreads), 0, st, gi.adj, s0, s1, s2, no_bits, d->B, d->N, rowmask, colmask,
                          molflag, nbr, cnbr));
  } else if (gi.compact()) {
    CU_TRY(launch_chain(adj_pack_kernel<true>, dim3(d->B), dim3(kThreads), smem, st, no_adj, 0ll, 0ll, 0ll, gi.adj_bits, d->N, NW, rowmask, colmask,
                        molflag, nbr, cnbr));
  } else {
    CU_TRY(launch_chain(adj_pack_kernel<false>, dim3(d->B), dim3(kThreads), smem, st, gi.adj, s0, s1, s2, no_bits, d->N, NW, rowmask, colmask,
                        molflag, nbr, cnbr));
  }
  LAUNCH_CHECK();
  return OCGCN_OK;
}

static int run_adj_pack(const ocgcn_desc* d, const Layout& lo, const GraphIn& gi, char* saved, cudaStream_t st) {
  return launch_adj_pack(d, gi, lo.NW, reinterpret_cast<u64*>(saved + lo.rowmask), reinterpret_cast<u64*>(saved + lo.colmask),
                         reinterpret_cast<unsigned*>(saved + lo.molflag), reinterpret_cast<uint4*>(saved + lo.nbr),
                         reinterpret_cast<uint4*>(saved + lo.cnbr), st);
}

static int run_bn_finalize(const ocgcn_desc* d, const Layout& lo, const ocgcn_params* p, int l, char* saved, char* scratch,
                           cudaStream_t st) {
  ProfScope prof(KK_BN_FINALIZE, st);
  BnFinalizeArgs f;
  const int F = d->F[l + 1];
  float* stats = reinterpret_cast<float*>(saved + lo.stats[l]);
  f.partials = reinterpret_cast<const float*>(scratch + lo.partials);
  f.T = lo.tiles; f.F = F; f.B = d->B; f.N = d->N; f.MT = lo.MT;
  f.training = d->training; f.eps = d->eps; f.momentum = d->momentum;
  f.gamma = p->gamma[l]; f.beta = p->beta[l];
  f.running_mean = p->running_mean[l]; f.running_var = p->running_var[l];
  f.nbt = reinterpret_cast<long long*>(p->num_batches_tracked[l]);
  f.mean = stats; f.invstd = stats + F; f.scale = stats + 2 * F; f.shift = stats + 3 * F;
  if (d->training && chan_wide(lo.tiles)) CU_TRY(launch_chain(bn_finalize_kernel<kChanColsWide>, dim3((F + kChanColsWide - 1) / kChanColsWide), dim3(kRedThreads), 0, st, f));
  else CU_TRY(launch_chain(bn_finalize_kernel<kChanCols>, dim3((F + kChanCols - 1) / kChanCols), dim3(kRedThreads), 0, st, f));
  LAUNCH_CHECK();
  return OCGCN_OK;
}

static int run_gemm_tn(const float* A, const float* Bm, size_t R, int M, int Nn, float* partial, cudaStream_t st, int* splits_out) {
  ProfScope prof(KK_GEMM_TN, st);
  int splits, rps;
  gemm_tn_config(R, M, Nn, &splits, &rps);
  dim3 grid(splits, ((M + 127) / 128) * ((Nn + 127) / 128));
  gemm_tn_kernel<<<grid, kThreads, 0, st>>>(A, Bm, (int)R, M, Nn, rps, partial);
  LAUNCH_CHECK();
  *splits_out = splits;
  return OCGCN_OK;
}

static void add_reduce(ReduceJobs& j, const float* partial, float* out, int P, int count, int matrix, int tr_cols) {
  if (!out || count <= 0) return;
  const int n = j.n;
  j.partial[n] = partial; j.out[n] = out; j.P[n] = P; j.count[n] = count; j.matrix[n] = matrix; j.tr_cols[n] = tr_cols;
  j.first_cta[n + 1] = j.first_cta[n] + (reduce_job_vec4(matrix, tr_cols, count) ? (count + 127) / 128 : matrix ? (count + 31) / 32 : (count + kChanCols - 1) / kChanCols);
  j.n = n + 1;
}
static int run_reduce(const ReduceJobs& j, cudaStream_t st) {
  if (j.n == 0) return OCGCN_OK;
  ProfScope prof(KK_REDUCE, st);
  CU_TRY(launch_chain(reduce_partials_kernel, dim3(j.first_cta[j.n]), dim3(kRedThreads), 0, st, j));
  LAUNCH_CHECK();
  return OCGCN_OK;
}

// every parameter / buffer pointer must be non-null DEVICE memory (a module left on the host is an error, not a fault)
static int validate_params(const ocgcn_desc* d, const ocgcn_params* p, bool layer_only) {
  if (!p) return OCGCN_EINVAL;
  for (int l = 0; l < d->L; ++l) {
    ST_TRY(check_device_ptr(p->W[l]));
    ST_TRY(check_device_ptr(p->gamma[l]));
    ST_TRY(check_device_ptr(p->beta[l]));
    ST_TRY(check_device_ptr(p->running_mean[l]));
    ST_TRY(check_device_ptr(p->running_var[l]));
    if (d->has_bias) ST_TRY(check_device_ptr(p->b[l]));
    if (p->num_batches_tracked[l]) ST_TRY(check_device_ptr(p->num_batches_tracked[l]));
  }
  if (!layer_only) {
    ST_TRY(check_device_ptr(p->Wd));
    ST_TRY(check_device_ptr(p->bd));
  }
  return OCGCN_OK;
}

// The forward of one graph-conv layer's GEMM stage (producer P_X for l==0, P_MID otherwise) + BN statistics.
static int run_layer_forward(const ocgcn_desc* d, const Layout& lo, const GraphIn& gi, const ocgcn_params* p, int l,
                             char* saved, char* scratch, cudaStream_t st) {
  TileArgs a = base_tile_args(d, lo, gi.adj, saved);
  a.K = d->F[l]; a.NC = d->F[l + 1];
  a.Bm = p->W[l];
  a.bias = d->has_bias ? p->b[l] : nullptr;
  a.out0 = reinterpret_cast<float*>(saved + lo.Z[l]);
  a.prod_out = reinterpret_cast<float*>(saved + lo.S[l]);
  a.partials = reinterpret_cast<float*>(scratch + lo.partials);
  if (lo.mm) {
    a.prod_out = nullptr;
    a.prod_img = saved + lo.S[l]; a.img_chunks = lo.KP[l] / 8;
    a.wimg = scratch + lo.Wf[l]; a.ncb = (d->F[l + 1] + 127) / 128;
  }
  if (l == 0) {
    a.in0 = gi.x; a.xs0 = d->x_stride[0]; a.xs1 = d->x_stride[1]; a.xs2 = d->x_stride[2];
    a.x_bits = gi.x_bits; a.xw = (d->F[0] + 31) / 32;
    ST_TRY((launch_tile<P_X, E_ZSTATS>(lo, a, st)));
  } else {
    const float* stats = reinterpret_cast<const float*>(saved + lo.stats[l - 1]);
    a.in0 = reinterpret_cast<const float*>(saved + lo.Z[l - 1]);
    a.vec0 = stats + 2 * d->F[l]; a.vec1 = stats + 3 * d->F[l];
    a.arg_out = reinterpret_cast<unsigned char*>(saved + lo.arg[l - 1]);
    ST_TRY((launch_tile<P_MID, E_ZSTATS>(lo, a, st)));
  }
  return run_bn_finalize(d, lo, p, l, saved, scratch, st);
}

// E_DY / E_ATG_DY epilogues: saved state of layer `lp`, whose dY the tile kernel produces
static void set_prev_layer(TileArgs& a, const ocgcn_desc* d, const Layout& lo, int lp, char* saved, float* partials) {
  const int F = d->F[lp + 1];
  const float* stats = reinterpret_cast<const float*>(saved + lo.stats[lp]);
  a.arg_in = reinterpret_cast<const unsigned char*>(saved + lo.arg[lp]);
  a.zprev = reinterpret_cast<const float*>(saved + lo.Z[lp]);
  a.pv_mean = stats; a.pv_invstd = stats + F; a.pv_scale = stats + 2 * F; a.pv_shift = stats + 3 * F;
  a.partials = partials;
}

// Backward of one graph-conv layer given dY in buffer `g_in` (layer API: plain; encoder: produced by the fused epilogue of
// the layer above, together with its partial sums).
// Produces dgamma/dbeta/dW/db, and dHprev (= dX) in `g_out` unless skip_dx.
static int run_layer_backward(const ocgcn_desc* d, const Layout& lo, const float* adj, const ocgcn_params* p, const ocgcn_grads* g,
                              int l, bool plain, float* g_in, float* g_dz, float* g_out, bool skip_dx, char* saved, char* scratch,
                              cudaStream_t st, ReduceJobs& rj) {
  const int Fin = d->F[l], Fout = d->F[l + 1];
  const float* stats = reinterpret_cast<const float*>(saved + lo.stats[l]);
  float* coef = reinterpret_cast<float*>(scratch + lo.coef);
  float* partials = reinterpret_cast<float*>(scratch + lo.partials);
  float* prodpart = reinterpret_cast<float*>(scratch + lo.prodpart[l]);
  const float* Z = reinterpret_cast<const float*>(saved + lo.Z[l]);

  BwdDyArgs y;
  memset(&y, 0, sizeof(y));
  y.B = d->B; y.N = d->N; y.NW = lo.NW; y.MT = lo.MT; y.F = Fout; y.plain = plain ? 1 : 0;
  y.adj = adj; y.as0 = d->adj_stride[0]; y.as1 = d->adj_stride[1]; y.as2 = d->adj_stride[2];
  y.colmask = reinterpret_cast<const u64*>(saved + lo.colmask);
  y.cnbr = reinterpret_cast<const uint4*>(saved + lo.cnbr);
  y.Z = Z;
  y.arg = plain ? nullptr : reinterpret_cast<const unsigned char*>(saved + lo.arg[l]);
  y.mean = stats; y.invstd = stats + Fout; y.scale = stats + 2 * Fout; y.shift = stats + 3 * Fout;
  y.dH = g_in; y.partials = partials;
  if (plain) ST_TRY(launch_bwd_dy(lo, y, st));   // encoder: dY and its partial sums come from the previous fused epilogue

  BwdFinalizeArgs f;
  f.partials = partials; f.T = lo.tiles; f.F = Fout; f.training = d->training; f.n = (double)lo.rows;
  f.gamma = p->gamma[l]; f.invstd = stats + Fout;
  f.dgamma = g->dgamma[l]; f.dbeta = g->dbeta[l];
  const int cstride = (lo.Fmax + 3) / 4 * 4;
  f.g = coef; f.c1 = coef + cstride; f.c2 = coef + 2 * cstride;
  {
    ProfScope prof(KK_BWD_FINALIZE, st);
    if (chan_wide(lo.tiles)) CU_TRY(launch_chain(bwd_finalize_kernel<kChanColsWide>, dim3((Fout + kChanColsWide - 1) / kChanColsWide), dim3(kRedThreads), 0, st, f));
    else CU_TRY(launch_chain(bwd_finalize_kernel<kChanCols>, dim3((Fout + kChanCols - 1) / kChanCols), dim3(kRedThreads), 0, st, f));
    LAUNCH_CHECK();
  }

  TileArgs a = base_tile_args(d, lo, adj, saved);
  a.K = Fout; a.NC = Fin;
  a.in0 = g_in; a.in1 = Z;
  a.vec0 = coef; a.vec1 = coef + cstride; a.vec2 = coef + 2 * cstride; a.vec3 = stats; a.vec4 = stats + Fout;
  a.prod_out = g_dz;
  a.prod_partials = (d->has_bias && g->db[l]) ? prodpart : nullptr;
  a.Bm = reinterpret_cast<const float*>(scratch + lo.WT[l]);
  a.out0 = g_out;
  a.skip_gemm = skip_dx ? 1 : 0;
  float* gp = reinterpret_cast<float*>(scratch + lo.gemm_partial[l]);
  if (lo.mm) {
    a.prod_out = nullptr;
    a.prod_img = scratch + lo.dZimg; a.img_chunks = lo.KP[l + 1] / 8;
    a.wimg = saved + lo.Wb[l]; a.ncb = (Fin + 127) / 128;
    if (lo.fuse_dw) {   // dW_l = S_l^T . dZ accumulated in TMEM inside this kernel: no dZ image, no separate GEMM
      a.prod_img = nullptr;
      a.dw_img = saved + lo.S[l]; a.dw_chunks = lo.KP[l] / 8; a.dw_M = Fin; a.dw_partial = gp;
    }
  }
  if (plain) {
    ST_TRY((launch_tile<P_DZ, E_ATG>(lo, a, st)));
  } else {
    if (l > 0) set_prev_layer(a, d, lo, l - 1, saved, partials);
    ST_TRY((launch_tile<P_DZ, E_ATG_DY>(lo, a, st)));
  }

  int splits = 0;
  if (lo.fuse_dw) splits = mm_grid(lo.tiles);
  else if (lo.mm) ST_TRY(run_dw_mma(lo, saved + lo.S[l], scratch + lo.dZimg, lo.KP[l], lo.KP[l + 1], Fin, Fout, gp, st, &splits));
  else ST_TRY(run_gemm_tn(reinterpret_cast<const float*>(saved + lo.S[l]), g_dz, lo.rows, Fin, Fout, gp, st, &splits));
  add_reduce(rj, gp, g->dW[l], splits, Fin * Fout, 1, 0);
  add_reduce(rj, prodpart, (d->has_bias ? g->db[l] : nullptr), lo.tiles, Fout, 0, 0);
  return OCGCN_OK;
}

static int run_transposes(const TransposeJobs& jobs, int n, cudaStream_t st) {
  if (n == 0) return OCGCN_OK;
  ProfScope prof(KK_TRANSPOSE, st);
  dim3 grid(32, n);
  transpose_kernel<<<grid, 256, 0, st>>>(jobs);
  LAUNCH_CHECK();
  return OCGCN_OK;
}

}  // namespace ocgcn

using namespace ocgcn;

extern "C" {

int ocgcn_abi_version(void) { return OCGCN_ABI_VERSION; }

int ocgcn_profile_enable(int on) {
  g_prof_on.store(on != 0);
  return OCGCN_OK;
}
int ocgcn_profile_kinds(void) { return KK_COUNT; }
const char* ocgcn_profile_kind_name(int kind) { return (kind >= 0 && kind < KK_COUNT) ? kKindNames[kind] : "?"; }
int ocgcn_profile_collect(float* ms_by_kind, int* launches_by_kind, int n_kinds) {
  if (!ms_by_kind || !launches_by_kind || n_kinds < KK_COUNT) return OCGCN_EINVAL;
  for (int i = 0; i < n_kinds; ++i) { ms_by_kind[i] = 0.f; launches_by_kind[i] = 0; }
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) {
    float ms = 0.f;
    CU_TRY(cudaEventSynchronize(r.e1));
    CU_TRY(cudaEventElapsedTime(&ms, r.e0, r.e1));
    ms_by_kind[r.kind] += ms;
    launches_by_kind[r.kind] += 1;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  return OCGCN_OK;
}
int ocgcn_last_launch_count(void) { return tl_launches; }
int ocgcn_last_cuda_error(void) { return tl_cuda_err; }

const char* ocgcn_status_string(int s) {
  switch (s) {
    case OCGCN_OK: return "ok";
    case OCGCN_EINVAL: return "ocgcn: invalid argument (shape, null/host pointer or enum)";
    case OCGCN_EUNSUPPORTED: return "ocgcn: unsupported size or mode (N <= 256, layers <= 16, B*N < 2^31)";
    case OCGCN_EALIGN: return "ocgcn: misaligned pointer";
    case OCGCN_EARCH: return "ocgcn: current CUDA device is not sm_100 (B200)";
    case OCGCN_ECUDA: return "ocgcn: CUDA error (see ocgcn_last_cuda_error)";
    default: return "ocgcn: unknown status";
  }
}

int ocgcn_saved_tensor(const ocgcn_desc* desc, int what, int layer, size_t* offset, size_t* bytes) {
  Layout lo;
  const bool layer_only = desc && desc->E <= 0;
  ST_TRY(make_layout(desc, layer_only, &lo));
  if (!offset || !bytes || layer < 0 || layer >= desc->L) return OCGCN_EINVAL;
  const size_t rows = lo.rows;
  switch (what) {
    case OCGCN_SAVED_Z: *offset = lo.Z[layer]; *bytes = rows * desc->F[layer + 1] * 4; break;
    case OCGCN_SAVED_ARGMAX: if (layer_only) return OCGCN_EINVAL; *offset = lo.arg[layer]; *bytes = rows * desc->F[layer + 1]; break;
    case OCGCN_SAVED_S: *offset = lo.S[layer]; *bytes = lo.mm ? (size_t)lo.tiles * lo.KP[layer] * 256 : rows * desc->F[layer] * 4; break;
    case OCGCN_SAVED_STATS: *offset = lo.stats[layer]; *bytes = (size_t)4 * desc->F[layer + 1] * 4; break;
    case OCGCN_SAVED_ROWMASK: *offset = lo.rowmask; *bytes = rows * lo.NW * 8; break;
    case OCGCN_SAVED_MOLFLAG: *offset = lo.molflag; *bytes = (size_t)desc->B * 4; break;
    default: return OCGCN_EINVAL;
  }
  return OCGCN_OK;
}

int ocgcn_query_workspace(const ocgcn_desc* desc, size_t* saved_bytes, size_t* scratch_bytes) {
  Layout lo;
  const bool layer_only = desc && desc->E <= 0;
  ST_TRY(make_layout(desc, layer_only, &lo));
  if (saved_bytes) *saved_bytes = lo.saved_total;
  if (scratch_bytes) *scratch_bytes = lo.scratch_total;
  return OCGCN_OK;
}

static int encoder_forward_impl(const ocgcn_desc* d, const GraphIn& gi, const ocgcn_params* p, void* saved_v, void* scratch_v, float* out,
                                void* stream) {
  tl_launches = 0;
  tl_tile_parity = 0;
  Layout lo;
  ST_TRY(make_layout(d, false, &lo));
  ST_TRY(validate_params(d, p, false));
  if (!saved_v || !scratch_v || !out) return OCGCN_EINVAL;
  ST_TRY(check_device());
  ST_TRY(check_graph_in(gi));
  ST_TRY(check_device_ptr(out));
  if (((uintptr_t)saved_v & 255) || ((uintptr_t)scratch_v & 255)) return OCGCN_EALIGN;
  char* saved = static_cast<char*>(saved_v);
  char* scratch = static_cast<char*>(scratch_v);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int L = d->L;

  ST_TRY(run_adj_pack(d, lo, gi, saved, st));
  if (lo.mm) {
    WPrepJobs wj;
    memset(&wj, 0, sizeof(wj));
    int nj = 0;
    for (int l = 0; l < L; ++l) add_wprep(wj, nj, p->W[l], scratch + lo.Wf[l], d->F[l], d->F[l + 1], 0, 2);
    add_wprep(wj, nj, p->Wd, scratch + lo.Wro, d->F[L], d->E, 1, 2);
    for (int l = 1; l < L; ++l) add_wprep(wj, nj, p->W[l], saved + lo.Wb[l], d->F[l + 1], d->F[l], 1, 2);   // for the backward call
    add_wprep(wj, nj, p->Wd, saved + lo.Wrob, d->E, d->F[L], 0, 2);
    ST_TRY(run_wprep(wj, nj, st));
  } else {
    TransposeJobs tj;
    memset(&tj, 0, sizeof(tj));
    tj.in[0] = p->Wd; tj.out[0] = reinterpret_cast<float*>(scratch + lo.WdT); tj.R[0] = d->E; tj.C[0] = d->F[L];
    ST_TRY(run_transposes(tj, 1, st));
  }
  // Training batches of at most one tile per SM (the Tox21 shape, C2): the whole forward in ONE cooperative kernel, accumulators parked in
  // TMEM across the per-layer grid barrier (ocgcn_fused.cuh); it writes the same saved state as the layer-by-layer kernels below.
  {
    bool fused = lo.mm && d->training && fused_train_enabled() && d->N <= 128 && L <= kFusedMaxLayers && d->E <= 128;
    for (int l = 0; l <= L && fused; ++l) fused = d->F[l] <= 128 && (l == 0 || d->F[l] % 4 == 0);
    const bool coop = fused && lo.tiles <= 148;                    // one tile per SM: all stages in one cooperative launch
    const bool per_layer = fused && !coop && fused_layer_enabled(); // larger batches: the same kernel, one persistent launch per stage
    if (coop || per_layer) {
      FusedTrainArgs t;
      memset(&t, 0, sizeof(t));
      FusedEvalArgs& a = t.e;
      a.B = d->B; a.N = d->N; a.MT = lo.MT; a.ntiles = lo.tiles; a.L = L; a.E = d->E; a.eps = d->eps;
      for (int l = 0; l <= L; ++l) { a.F[l] = d->F[l]; t.KP[l] = lo.KP[l]; }
      a.x = gi.x; a.xs0 = d->x_stride[0]; a.xs1 = d->x_stride[1]; a.xs2 = d->x_stride[2];
      a.adj = gi.adj; a.as0 = d->adj_stride[0]; a.as1 = d->adj_stride[1]; a.as2 = d->adj_stride[2];
      a.x_bits = gi.x_bits; a.adj_bits = gi.adj_bits; a.xw = (d->F[0] + 31) / 32; a.aw = (d->N + 31) / 32;
      a.nbr = reinterpret_cast<const uint4*>(saved + lo.nbr);
      for (int l = 0; l < L; ++l) {
        a.wimg[l] = scratch + lo.Wf[l];
        a.bias[l] = d->has_bias ? p->b[l] : nullptr;
        a.gamma[l] = p->gamma[l]; a.beta[l] = p->beta[l];
        a.running_mean[l] = p->running_mean[l]; a.running_var[l] = p->running_var[l];
        t.rm[l] = p->running_mean[l]; t.rv[l] = p->running_var[l];
        t.nbt[l] = reinterpret_cast<long long*>(p->num_batches_tracked[l]);
        t.Z[l] = reinterpret_cast<float*>(saved + lo.Z[l]);
        t.arg[l] = reinterpret_cast<unsigned char*>(saved + lo.arg[l]);
        t.S[l] = saved + lo.S[l];
        t.stats[l] = reinterpret_cast<float*>(saved + lo.stats[l]);
      }
      a.wimg[L] = scratch + lo.Wro;
      a.bias[L] = p->bd;
      a.out = out;
      t.S[L] = saved + lo.HL;
      t.D = reinterpret_cast<float*>(saved + lo.D);
      t.partials = reinterpret_cast<float*>(scratch + lo.partials);
      t.momentum = d->momentum;
      const size_t smem = fused_train_smem_bytes();
      ST_TRY((ensure_dynamic_smem<encoder_train_fused_kernel>(smem)));
      if (coop) {
        t.l_begin = 0; t.l_end = L + 1;
        ProfScope prof(KK_FUSED_TRAIN, st);
        CU_TRY(launch_cooperative(encoder_train_fused_kernel, dim3(lo.tiles), dim3(kFusedThreads), smem, st, t));
        LAUNCH_CHECK();
      } else {
        for (int l = 0; l <= L; ++l) {
          t.l_begin = l; t.l_end = l + 1;
          {
            ProfScope prof(l == 0 ? KK_FWD_FIRST : l < L ? KK_FWD_MID : KK_FWD_READOUT, st);
            CU_TRY(launch_chain(encoder_train_fused_kernel, dim3(148), dim3(kFusedThreads), smem, st, t));
            LAUNCH_CHECK();
          }
          if (l < L) ST_TRY(run_bn_finalize(d, lo, p, l, saved, scratch, st));
        }
      }
      CU_TRY(cudaMemcpyAsync(saved + lo.O, out, (size_t)d->B * d->E * 4, cudaMemcpyDeviceToDevice, st));
      return OCGCN_OK;
    }
  }
  for (int l = 0; l < L; ++l) ST_TRY(run_layer_forward(d, lo, gi, p, l, saved, scratch, st));

  // readout: H_L = pool(tanh(BN(Z_L))), D = tanh(H_L Wd^T + bd), O = tanh(sum_i D)
  TileArgs a = base_tile_args(d, lo, gi.adj, saved);
  const float* stats = reinterpret_cast<const float*>(saved + lo.stats[L - 1]);
  a.K = d->F[L]; a.NC = d->E;
  a.in0 = reinterpret_cast<const float*>(saved + lo.Z[L - 1]);
  a.vec0 = stats + 2 * d->F[L]; a.vec1 = stats + 3 * d->F[L];
  a.arg_out = reinterpret_cast<unsigned char*>(saved + lo.arg[L - 1]);
  a.prod_out = reinterpret_cast<float*>(saved + lo.HL);
  a.Bm = reinterpret_cast<const float*>(scratch + lo.WdT);
  a.bias = p->bd;
  a.out0 = reinterpret_cast<float*>(saved + lo.D);
  a.out1 = out;
  if (lo.mm) {
    a.prod_out = nullptr;
    a.prod_img = saved + lo.HL; a.img_chunks = lo.KP[L] / 8;
    a.wimg = scratch + lo.Wro; a.ncb = (d->E + 127) / 128;
  }
  ST_TRY((launch_tile<P_RO, E_READOUT>(lo, a, st)));
  CU_TRY(cudaMemcpyAsync(saved + lo.O, out, (size_t)d->B * d->E * 4, cudaMemcpyDeviceToDevice, st));
  return OCGCN_OK;
}

int ocgcn_encoder_forward(const ocgcn_desc* d, const float* x, const float* adj, const ocgcn_params* p, void* saved_v,
                          void* scratch_v, float* out, void* stream) {
  const GraphIn gi = {x, adj, nullptr, nullptr};
  return encoder_forward_impl(d, gi, p, saved_v, scratch_v, out, stream);
}

int ocgcn_encoder_forward_compact(const ocgcn_desc* d, const uint32_t* x_bits, const uint32_t* adj_bits, const ocgcn_params* p,
                                  void* saved_v, void* scratch_v, float* out, void* stream) {
  if (!x_bits || !adj_bits) return OCGCN_EINVAL;
  const GraphIn gi = {nullptr, nullptr, x_bits, adj_bits};
  return encoder_forward_impl(d, gi, p, saved_v, scratch_v, out, stream);
}

// ---- single-kernel eval forward (ocgcn_fused.cuh) --------------------------------------------------------------------------
struct EvalLayout {
  size_t rowmask, colmask, molflag, nbr, cnbr, Wf[OCGCN_MAX_LAYERS + 1], total;
  int NW, MT, tiles;
};
// the fused eval kernel serves the tensor-core classes in eval mode when every width fits one 128-column block
static int make_eval_layout(const ocgcn_desc* d, EvalLayout* lo) {
  memset(lo, 0, sizeof(*lo));
  if (!d || d->B <= 0 || d->N <= 0 || d->L <= 0 || d->E <= 0) return OCGCN_EINVAL;
  if (d->mode != OCGCN_TF32 && d->mode != OCGCN_BF16) return OCGCN_EUNSUPPORTED;
  if (d->training || d->N > 128 || d->L > kFusedMaxLayers || d->E > 128) return OCGCN_EUNSUPPORTED;
  for (int l = 0; l <= d->L; ++l) {
    if (d->F[l] <= 0) return OCGCN_EINVAL;
    if (d->F[l] > 128) return OCGCN_EUNSUPPORTED;
  }
  if ((size_t)d->B * d->N >= (size_t)1 << 31) return OCGCN_EUNSUPPORTED;
  lo->NW = (d->N + 63) / 64;
  lo->MT = 128 / d->N;
  lo->tiles = (d->B + lo->MT - 1) / lo->MT;
  const size_t rows = (size_t)d->B * d->N;
  size_t off = 0;
  auto take = [&off](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
  lo->rowmask = take(rows * lo->NW * 8);
  lo->colmask = take(rows * lo->NW * 8);
  lo->molflag = take((size_t)d->B * 4);
  lo->nbr = take(rows * 16);
  lo->cnbr = take(rows * 16);
  for (int l = 0; l < d->L; ++l) lo->Wf[l] = take(wimg_bytes(d->F[l], d->F[l + 1], 2));
  lo->Wf[d->L] = take(wimg_bytes(d->F[d->L], d->E, 2));
  lo->total = off;
  return OCGCN_OK;
}

int ocgcn_query_eval_workspace(const ocgcn_desc* desc, size_t* bytes) {
  EvalLayout lo;
  ST_TRY(make_eval_layout(desc, &lo));
  if (bytes) *bytes = lo.total;
  return OCGCN_OK;
}

static int encoder_forward_eval_impl(const ocgcn_desc* d, const GraphIn& gi, const ocgcn_params* p, void* workspace_v, float* out,
                                     void* stream) {
  tl_launches = 0;
  tl_tile_parity = 0;
  EvalLayout lo;
  ST_TRY(make_eval_layout(d, &lo));
  ST_TRY(validate_params(d, p, false));
  if (!workspace_v || !out) return OCGCN_EINVAL;
  ST_TRY(check_device());
  ST_TRY(check_graph_in(gi));
  ST_TRY(check_device_ptr(out));
  if ((uintptr_t)workspace_v & 255) return OCGCN_EALIGN;
  char* ws = static_cast<char*>(workspace_v);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int L = d->L;
  ST_TRY(launch_adj_pack(d, gi, lo.NW, reinterpret_cast<u64*>(ws + lo.rowmask), reinterpret_cast<u64*>(ws + lo.colmask),
                         reinterpret_cast<unsigned*>(ws + lo.molflag), reinterpret_cast<uint4*>(ws + lo.nbr),
                         reinterpret_cast<uint4*>(ws + lo.cnbr), st));
  WPrepJobs wj;
  memset(&wj, 0, sizeof(wj));
  int nj = 0;
  for (int l = 0; l < L; ++l) add_wprep(wj, nj, p->W[l], ws + lo.Wf[l], d->F[l], d->F[l + 1], 0, 2);
  add_wprep(wj, nj, p->Wd, ws + lo.Wf[L], d->F[L], d->E, 1, 2);
  ST_TRY(run_wprep(wj, nj, st));

  FusedEvalArgs a;
  memset(&a, 0, sizeof(a));
  a.B = d->B; a.N = d->N; a.MT = lo.MT; a.ntiles = lo.tiles; a.L = L; a.E = d->E; a.eps = d->eps;
  for (int l = 0; l <= L; ++l) a.F[l] = d->F[l];
  a.x = gi.x; a.xs0 = d->x_stride[0]; a.xs1 = d->x_stride[1]; a.xs2 = d->x_stride[2];
  a.adj = gi.adj; a.as0 = d->adj_stride[0]; a.as1 = d->adj_stride[1]; a.as2 = d->adj_stride[2];
  a.x_bits = gi.x_bits; a.adj_bits = gi.adj_bits; a.xw = (d->F[0] + 31) / 32; a.aw = (d->N + 31) / 32;
  a.nbr = reinterpret_cast<const uint4*>(ws + lo.nbr);
  for (int l = 0; l < L; ++l) {
    a.wimg[l] = ws + lo.Wf[l];
    a.bias[l] = d->has_bias ? p->b[l] : nullptr;
    a.gamma[l] = p->gamma[l]; a.beta[l] = p->beta[l];
    a.running_mean[l] = p->running_mean[l]; a.running_var[l] = p->running_var[l];
  }
  a.wimg[L] = ws + lo.Wf[L];
  a.bias[L] = p->bd;
  a.out = out;
  const size_t smem = fused_eval_smem_bytes();
  ST_TRY((ensure_dynamic_smem<encoder_eval_fused_kernel>(smem)));
  ProfScope prof(KK_FUSED_EVAL, st);
  CU_TRY(launch_chain(encoder_eval_fused_kernel, dim3(lo.tiles < 148 ? lo.tiles : 148), dim3(kFusedThreads), smem, st, a));
  LAUNCH_CHECK();
  return OCGCN_OK;
}

int ocgcn_encoder_forward_eval(const ocgcn_desc* d, const float* x, const float* adj, const ocgcn_params* p, void* workspace_v,
                               float* out, void* stream) {
  const GraphIn gi = {x, adj, nullptr, nullptr};
  return encoder_forward_eval_impl(d, gi, p, workspace_v, out, stream);
}

int ocgcn_encoder_forward_eval_compact(const ocgcn_desc* d, const uint32_t* x_bits, const uint32_t* adj_bits, const ocgcn_params* p,
                                       void* workspace_v, float* out, void* stream) {
  if (!x_bits || !adj_bits) return OCGCN_EINVAL;
  const GraphIn gi = {nullptr, nullptr, x_bits, adj_bits};
  return encoder_forward_eval_impl(d, gi, p, workspace_v, out, stream);
}

// adj == nullptr: the forward call was ocgcn_encoder_forward_compact (every molecule's adjacency is exactly 0/1, so the kernels walk the
// packed records / masks in `saved` and never touch the dense tensor)
static int encoder_backward_impl(const ocgcn_desc* d, const float* grad_out, const float* adj, const ocgcn_params* p,
                                 const ocgcn_grads* g, void* saved_v, void* scratch_v, void* stream) {
  tl_launches = 0;
  tl_tile_parity = d ? ((d->L + 1) & 1) : 0;   // the forward's L + 1 tile launches ended on parity L & 1: start on the other direction
  Layout lo;
  ST_TRY(make_layout(d, false, &lo));
  ST_TRY(validate_params(d, p, false));
  if (!g || !saved_v || !scratch_v || !g->dWd || !g->dbd) return OCGCN_EINVAL;
  for (int l = 0; l < d->L; ++l)
    if (!g->dW[l] || !g->dgamma[l] || !g->dbeta[l]) return OCGCN_EINVAL;
  ST_TRY(check_device());
  ST_TRY(check_device_ptr(grad_out));
  if (adj) ST_TRY(check_device_ptr(adj));
  if (((uintptr_t)saved_v & 255) || ((uintptr_t)scratch_v & 255)) return OCGCN_EALIGN;
  char* saved = static_cast<char*>(saved_v);
  char* scratch = static_cast<char*>(scratch_v);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int L = d->L;

  if (lo.mm) {
    // (weight operand images for the backward products were written into `saved` by the forward call)
  } else {
    TransposeJobs tj;
    memset(&tj, 0, sizeof(tj));
    int nj = 0;
    for (int l = 1; l < L; ++l) {
      tj.in[nj] = p->W[l]; tj.out[nj] = reinterpret_cast<float*>(scratch + lo.WT[l]); tj.R[nj] = d->F[l]; tj.C[nj] = d->F[l + 1];
      ++nj;
    }
    ST_TRY(run_transposes(tj, nj, st));
  }

  float* G[3] = {reinterpret_cast<float*>(scratch + lo.G[0]), reinterpret_cast<float*>(scratch + lo.G[1]),
                 reinterpret_cast<float*>(scratch + lo.G[2])};
  float* prodpart = reinterpret_cast<float*>(scratch + lo.prodpart[L]);
  float* dU = reinterpret_cast<float*>(scratch + lo.dU);
  float* gp = reinterpret_cast<float*>(scratch + lo.gemm_partial[L]);
  ReduceJobs rj;
  memset(&rj, 0, sizeof(rj));

  // readout backward: dU = dO(1-O^2)(1-D^2); dH_L = dU.Wd; dWd = dU^T.H_L; dbd = sum dU
  TileArgs a = base_tile_args(d, lo, adj, saved);
  a.K = d->E; a.NC = d->F[L];
  a.in0 = reinterpret_cast<const float*>(saved + lo.D);
  a.in1 = grad_out;
  a.in2 = reinterpret_cast<const float*>(saved + lo.O);
  a.prod_out = dU;
  a.prod_partials = prodpart;
  a.Bm = p->Wd;
  a.out0 = G[0];
  if (lo.mm) {
    a.prod_out = nullptr;
    a.prod_img = scratch + lo.dUimg; a.img_chunks = lo.KPE / 8;
    a.wimg = saved + lo.Wrob; a.ncb = (d->F[L] + 127) / 128;
    if (lo.fuse_dw) {   // dWd^T = H_L^T . dU accumulated inside this kernel (transposed back by the reduction)
      a.prod_img = nullptr;
      a.dw_img = saved + lo.HL; a.dw_chunks = lo.KP[L] / 8; a.dw_M = d->F[L]; a.dw_partial = gp;
    }
  }
  set_prev_layer(a, d, lo, L - 1, saved, reinterpret_cast<float*>(scratch + lo.partials));
  ST_TRY((launch_tile<P_DU, E_DY>(lo, a, st)));
  int splits = 0;
  if (lo.fuse_dw) splits = mm_grid(lo.tiles);
  else if (lo.mm) ST_TRY(run_dw_mma(lo, scratch + lo.dUimg, saved + lo.HL, lo.KPE, lo.KP[L], d->E, d->F[L], gp, st, &splits));
  else ST_TRY(run_gemm_tn(dU, reinterpret_cast<const float*>(saved + lo.HL), lo.rows, d->E, d->F[L], gp, st, &splits));
  add_reduce(rj, gp, g->dWd, splits, d->E * d->F[L], 1, lo.fuse_dw ? d->F[L] : 0);
  add_reduce(rj, prodpart, g->dbd, lo.tiles, d->E, 0, 0);

  int cur = 0;
  for (int l = L - 1; l >= 0; --l) {
    const int dz = (cur + 1) % 3, nxt = (cur + 2) % 3;
    ST_TRY(run_layer_backward(d, lo, adj, p, g, l, false, G[cur], G[dz], G[nxt], l == 0, saved, scratch, st, rj));
    cur = nxt;
  }
  return run_reduce(rj, st);   // every dW_l, db_l, dWd, dbd in one fixed-order reduction launch
}

int ocgcn_encoder_backward(const ocgcn_desc* d, const float* grad_out, const float* x, const float* adj, const ocgcn_params* p,
                           const ocgcn_grads* g, void* saved_v, void* scratch_v, void* stream) {
  (void)x;
  if (!adj) return OCGCN_EINVAL;
  return encoder_backward_impl(d, grad_out, adj, p, g, saved_v, scratch_v, stream);
}

int ocgcn_encoder_backward_compact(const ocgcn_desc* d, const float* grad_out, const ocgcn_params* p, const ocgcn_grads* g, void* saved_v,
                                   void* scratch_v, void* stream) {
  return encoder_backward_impl(d, grad_out, nullptr, p, g, saved_v, scratch_v, stream);
}

int ocgcn_layer_forward(const ocgcn_desc* d, const float* x, const float* adj, const ocgcn_params* p, void* saved_v, void* scratch_v,
                        float* y, void* stream) {
  tl_launches = 0;
  tl_tile_parity = 0;
  Layout lo;
  ST_TRY(make_layout(d, true, &lo));
  ST_TRY(validate_params(d, p, true));
  if (!saved_v || !scratch_v || !y) return OCGCN_EINVAL;
  ST_TRY(check_device());
  ST_TRY(check_device_ptr(x));
  ST_TRY(check_device_ptr(adj));
  ST_TRY(check_device_ptr(y));
  const GraphIn gi = {x, adj, nullptr, nullptr};
  if (((uintptr_t)saved_v & 255) || ((uintptr_t)scratch_v & 255)) return OCGCN_EALIGN;
  char* saved = static_cast<char*>(saved_v);
  char* scratch = static_cast<char*>(scratch_v);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ST_TRY(run_adj_pack(d, lo, gi, saved, st));
  if (lo.mm) {
    WPrepJobs wj;
    memset(&wj, 0, sizeof(wj));
    int nj = 0;
    add_wprep(wj, nj, p->W[0], scratch + lo.Wf[0], d->F[0], d->F[1], 0, 2);
    if (d->need_dx) add_wprep(wj, nj, p->W[0], saved + lo.Wb[0], d->F[1], d->F[0], 1, 2);
    ST_TRY(run_wprep(wj, nj, st));
  }
  ST_TRY(run_layer_forward(d, lo, gi, p, 0, saved, scratch, st));
  const int F = d->F[1];
  const float* stats = reinterpret_cast<const float*>(saved + lo.stats[0]);
  const size_t total = lo.rows * F;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  {
    ProfScope prof(KK_BN_APPLY, st);
    bn_apply_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(saved + lo.Z[0]), stats + 2 * F, stats + 3 * F, y, total, F);
    LAUNCH_CHECK();
  }
  return OCGCN_OK;
}

int ocgcn_layer_backward(const ocgcn_desc* d, const float* grad_y, const float* x, const float* adj, const ocgcn_params* p,
                         const ocgcn_grads* g, float* grad_x, void* saved_v, void* scratch_v, void* stream) {
  (void)x;
  tl_launches = 0;
  tl_tile_parity = 0;
  Layout lo;
  ST_TRY(make_layout(d, true, &lo));
  ST_TRY(validate_params(d, p, true));
  if (!g || !saved_v || !scratch_v || !g->dW[0] || !g->dgamma[0] || !g->dbeta[0]) return OCGCN_EINVAL;
  if (d->need_dx && !grad_x) return OCGCN_EINVAL;
  ST_TRY(check_device());
  ST_TRY(check_device_ptr(grad_y));
  ST_TRY(check_device_ptr(adj));
  if (((uintptr_t)saved_v & 255) || ((uintptr_t)scratch_v & 255)) return OCGCN_EALIGN;
  char* saved = static_cast<char*>(saved_v);
  char* scratch = static_cast<char*>(scratch_v);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d->need_dx && lo.mm) {
    // (W^T operand image written by the forward call)
  } else if (d->need_dx) {
    TransposeJobs tj;
    memset(&tj, 0, sizeof(tj));
    tj.in[0] = p->W[0]; tj.out[0] = reinterpret_cast<float*>(scratch + lo.WT[0]); tj.R[0] = d->F[0]; tj.C[0] = d->F[1];
    ST_TRY(run_transposes(tj, 1, st));
  }
  float* dz = reinterpret_cast<float*>(scratch + lo.G[0]);
  ReduceJobs rj;
  memset(&rj, 0, sizeof(rj));
  ST_TRY(run_layer_backward(d, lo, adj, p, g, 0, true, const_cast<float*>(grad_y), dz, grad_x, !d->need_dx, saved, scratch, st, rj));
  return run_reduce(rj, st);
}

static int bits_grid(long long work_items) {
  long long g = (work_items + 255) / 256;
  const long long cap = 148 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int ocgcn_expand_compact(int B, int N, int F0, const uint32_t* x_bits, const uint32_t* adj_bits, const int32_t* index, int n_src,
                         int32_t* index_oob, float* x, float* adj, void* stream) {
  tl_launches = 0;
  tl_tile_parity = 0;
  if (B <= 0 || N <= 0 || F0 <= 0) return OCGCN_EINVAL;
  if ((x != nullptr) != (x_bits != nullptr) || (adj != nullptr) != (adj_bits != nullptr) || (!x && !adj)) return OCGCN_EINVAL;
  if (index && (n_src <= 0 || !index_oob)) return OCGCN_EINVAL;
  ST_TRY(check_device());
  if (x) { ST_TRY(check_device_ptr(x)); ST_TRY(check_device_ptr(x_bits)); }
  if (adj) { ST_TRY(check_device_ptr(adj)); ST_TRY(check_device_ptr(adj_bits)); }
  if (index) { ST_TRY(check_device_ptr(index)); ST_TRY(check_device_ptr(index_oob)); }
  if (((uintptr_t)x & 15) || ((uintptr_t)adj & 15)) return OCGCN_EALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (index) CU_TRY(cudaMemsetAsync(index_oob, 0, sizeof(int32_t), st));
  const long long rows = (long long)B * N;
  if (adj) {
    ProfScope ps(KK_EXPAND, st);
    expand_bits_kernel<<<bits_grid(rows * ((N + 3) / 4)), 256, 0, st>>>(adj_bits, index, n_src, index_oob, N, N, (N + 31) / 32, rows, adj);
    LAUNCH_CHECK();
  }
  if (x) {
    ProfScope ps(KK_EXPAND, st);
    expand_bits_kernel<<<bits_grid(rows * ((F0 + 3) / 4)), 256, 0, st>>>(x_bits, index, n_src, index_oob, N, F0, (F0 + 31) / 32, rows, x);
    LAUNCH_CHECK();
  }
  return OCGCN_OK;
}

static int pool_args(int B, int N, int D, const float* y, const float* adj, const int64_t* adj_stride, PoolArgs* a) {
  if (B <= 0 || N <= 0 || D <= 0 || !adj_stride) return OCGCN_EINVAL;
  if (N > OCGCN_MAX_ATOMS) return OCGCN_EUNSUPPORTED;
  ST_TRY(check_device());
  ST_TRY(check_device_ptr(y));
  ST_TRY(check_device_ptr(adj));
  memset(a, 0, sizeof(*a));
  a->B = B; a->N = N; a->D = D; a->y = y; a->adj = adj;
  a->as0 = adj_stride[0]; a->as1 = adj_stride[1]; a->as2 = adj_stride[2];
  return OCGCN_OK;
}

int ocgcn_pool_forward(int B, int N, int D, const float* y, const float* adj, const int64_t* adj_stride, float* pooled, uint8_t* argmax,
                       void* stream) {
  tl_launches = 0;
  tl_tile_parity = 0;
  PoolArgs a;
  ST_TRY(pool_args(B, N, D, y, adj, adj_stride, &a));
  ST_TRY(check_device_ptr(pooled));
  if (!argmax) return OCGCN_EINVAL;
  a.pooled = pooled; a.arg = argmax;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = pool_smem_bytes(N, false);
  ST_TRY((ensure_dynamic_smem<pool_fwd_kernel>(smem)));
  ProfScope ps(KK_POOL_FWD, st);
  CU_TRY(launch_chain(pool_fwd_kernel, dim3((D + kPoolCols - 1) / kPoolCols, B), dim3(kPoolCols), smem, st, a));
  LAUNCH_CHECK();
  return OCGCN_OK;
}

int ocgcn_pool_backward(int B, int N, int D, const float* y, const float* adj, const int64_t* adj_stride, const uint8_t* argmax,
                        const float* grad_pooled, float* grad_y, void* stream) {
  tl_launches = 0;
  tl_tile_parity = 0;
  PoolArgs a;
  ST_TRY(pool_args(B, N, D, y, adj, adj_stride, &a));
  ST_TRY(check_device_ptr(grad_pooled));
  ST_TRY(check_device_ptr(grad_y));
  if (!argmax) return OCGCN_EINVAL;
  a.arg = const_cast<uint8_t*>(argmax); a.grad_pooled = grad_pooled; a.grad_y = grad_y;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = pool_smem_bytes(N, true);
  ST_TRY((ensure_dynamic_smem<pool_bwd_kernel>(smem)));
  ProfScope ps(KK_POOL_BWD, st);
  CU_TRY(launch_chain(pool_bwd_kernel, dim3((D + kPoolCols - 1) / kPoolCols, B), dim3(kPoolCols), smem, st, a));
  LAUNCH_CHECK();
  return OCGCN_OK;
}

int ocgcn_pack_compact(int B, int N, int F0, const float* x, const int64_t* x_stride, const float* adj, const int64_t* adj_stride,
                       uint32_t* x_bits, uint32_t* adj_bits, int32_t* nonbinary, void* stream) {
  tl_launches = 0;
  tl_tile_parity = 0;
  if (B <= 0 || N <= 0 || F0 <= 0 || !nonbinary) return OCGCN_EINVAL;
  if ((x != nullptr) != (x_bits != nullptr) || (adj != nullptr) != (adj_bits != nullptr) || (!x && !adj)) return OCGCN_EINVAL;
  if ((x && !x_stride) || (adj && !adj_stride)) return OCGCN_EINVAL;
  ST_TRY(check_device());
  ST_TRY(check_device_ptr(nonbinary));
  if (x) { ST_TRY(check_device_ptr(x)); ST_TRY(check_device_ptr(x_bits)); }
  if (adj) { ST_TRY(check_device_ptr(adj)); ST_TRY(check_device_ptr(adj_bits)); }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU_TRY(cudaMemsetAsync(nonbinary, 0, sizeof(int32_t), st));
  const long long rows = (long long)B * N;
  if (adj) {
    ProfScope ps(KK_PACK, st);
    const int W = (N + 31) / 32;
    pack_bits_kernel<<<bits_grid(rows * W * 32), 256, 0, st>>>(adj, adj_stride[0], adj_stride[1], adj_stride[2], N, N, W, rows, adj_bits,
                                                              nonbinary);
    LAUNCH_CHECK();
  }
  if (x) {
    ProfScope ps(KK_PACK, st);
    const int W = (F0 + 31) / 32;
    pack_bits_kernel<<<bits_grid(rows * W * 32), 256, 0, st>>>(x, x_stride[0], x_stride[1], x_stride[2], N, F0, W, rows, x_bits, nonbinary);
    LAUNCH_CHECK();
  }
  return OCGCN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// MLP head + loss
// ---------------------------------------------------------------------------------------------------------
namespace {
struct HeadLayout {
  int L, ntiles, grid, T;
  // saved
  size_t U[OCGCN_HEAD_MAX_LAYERS], stat[OCGCN_HEAD_MAX_LAYERS], ncount, saved_total;
  // scratch
  size_t part[OCGCN_HEAD_MAX_LAYERS], loss_part, dY[OCGCN_HEAD_MAX_LAYERS], dpart[OCGCN_HEAD_MAX_LAYERS], dWpart[OCGCN_HEAD_MAX_LAYERS],
      dbpart[OCGCN_HEAD_MAX_LAYERS], scratch_total;
};
int head_layout(const ocgcn_head_desc* d, HeadLayout* lo) {
  memset(lo, 0, sizeof(*lo));
  if (!d || d->B <= 0 || d->n_layers <= 0 || d->n_layers > OCGCN_HEAD_MAX_LAYERS) return OCGCN_EINVAL;
  for (int i = 0; i <= d->n_layers; ++i) {
    if (d->width[i] <= 0) return OCGCN_EINVAL;
    if (d->width[i] > OCGCN_HEAD_MAX_WIDTH) return OCGCN_EUNSUPPORTED;
  }
  for (int i = 0; i < d->n_layers; ++i)
    if (d->act[i] < OCGCN_ACT_IDENTITY || d->act[i] > OCGCN_ACT_TANH) return OCGCN_EINVAL;
  if (d->loss < OCGCN_LOSS_NONE || d->loss > OCGCN_LOSS_MULTITASK) return OCGCN_EINVAL;
  if (d->training && d->n_layers > 1 && d->B < 2) return OCGCN_EINVAL;   // BatchNorm1d needs more than one value per channel
  const int L = d->n_layers;
  lo->L = L;
  lo->T = d->width[L];
  lo->ntiles = (d->B + HR - 1) / HR;
  lo->grid = lo->ntiles < kMaxPersistent ? lo->ntiles : kMaxPersistent;
  size_t off = 0;
  for (int i = 0; i + 1 < L; ++i) {
    lo->U[i] = off; off = align_up(off + (size_t)d->B * d->width[i + 1] * 4);
    lo->stat[i] = off; off = align_up(off + (size_t)2 * d->width[i + 1] * 4);
  }
  lo->ncount = off; off = align_up(off + (size_t)lo->T * 4);
  lo->saved_total = off;
  off = 0;
  for (int i = 0; i + 1 < L; ++i) {
    lo->part[i] = off; off = align_up(off + (size_t)lo->ntiles * 2 * d->width[i + 1] * 4);
    lo->dY[i] = off; off = align_up(off + (size_t)d->B * d->width[i + 1] * 4);
    lo->dpart[i] = off; off = align_up(off + (size_t)lo->ntiles * 2 * d->width[i + 1] * 4);
  }
  lo->loss_part = off; off = align_up(off + (size_t)lo->ntiles * 2 * lo->T * 4);
  for (int i = 0; i < L; ++i) {
    lo->dWpart[i] = off; off = align_up(off + (size_t)lo->grid * d->width[i] * d->width[i + 1] * 4);
    lo->dbpart[i] = off; off = align_up(off + (size_t)lo->grid * d->width[i + 1] * 4);
  }
  lo->scratch_total = off;
  return OCGCN_OK;
}
int head_validate(const ocgcn_head_desc* d, const ocgcn_head_params* p) {
  if (!p) return OCGCN_EINVAL;
  for (int i = 0; i < d->n_layers; ++i) {
    if (!p->W[i]) return OCGCN_EINVAL;
    if (i + 1 < d->n_layers && (!p->gamma[i] || !p->beta[i] || !p->running_mean[i] || !p->running_var[i])) return OCGCN_EINVAL;
  }
  return OCGCN_OK;
}
HeadBN head_bn(const ocgcn_head_desc* d, const ocgcn_head_params* p, const HeadLayout& lo, int i, char* saved, char* scratch) {
  HeadBN bn;
  bn.part = reinterpret_cast<const float*>(scratch + lo.part[i]);
  bn.gamma = p->gamma[i]; bn.beta = p->beta[i];
  bn.running_mean = p->running_mean[i]; bn.running_var = p->running_var[i];
  bn.nbt = reinterpret_cast<long long*>(p->num_batches_tracked[i]);
  bn.stat = reinterpret_cast<float*>(saved + lo.stat[i]);
  bn.act = d->act[i];
  return bn;
}
}  // namespace

int ocgcn_head_query_workspace(const ocgcn_head_desc* desc, size_t* saved_bytes, size_t* scratch_bytes) {
  HeadLayout lo;
  ST_TRY(head_layout(desc, &lo));
  if (saved_bytes) *saved_bytes = lo.saved_total;
  if (scratch_bytes) *scratch_bytes = lo.scratch_total;
  return OCGCN_OK;
}

int ocgcn_head_forward(const ocgcn_head_desc* d, const float* inp, const ocgcn_head_params* p, const float* target, void* saved_v,
                       void* scratch_v, float* pred, float* loss, void* stream) {
  tl_launches = 0;
  tl_tile_parity = 0;
  HeadLayout lo;
  ST_TRY(head_layout(d, &lo));
  ST_TRY(head_validate(d, p));
  if (!saved_v || !scratch_v || !pred) return OCGCN_EINVAL;
  ST_TRY(check_device());
  ST_TRY(check_device_ptr(inp));
  ST_TRY(check_device_ptr(pred));
  if (d->loss != OCGCN_LOSS_NONE) {
    ST_TRY(check_device_ptr(target));
    ST_TRY(check_device_ptr(loss));
  }
  if (((uintptr_t)saved_v & 255) || ((uintptr_t)scratch_v & 255)) return OCGCN_EALIGN;
  char* saved = static_cast<char*>(saved_v);
  char* scratch = static_cast<char*>(scratch_v);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int L = lo.L;
  for (int i = 0; i < L; ++i) {
    HeadFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.Kin = d->width[i]; a.Nout = d->width[i + 1];
    a.in = i == 0 ? inp : reinterpret_cast<const float*>(saved + lo.U[i - 1]);
    a.prev_bn = i > 0;
    if (i > 0) a.bn = head_bn(d, p, lo, i - 1, saved, scratch);
    a.W = p->W[i]; a.b = p->b[i];
    a.has_bn = i + 1 < L;
    a.out = a.has_bn ? reinterpret_cast<float*>(saved + lo.U[i]) : pred;
    a.part_out = a.has_bn ? reinterpret_cast<float*>(scratch + lo.part[i]) : nullptr;
    a.act = d->act[i];
    a.B = d->B; a.ntiles = lo.ntiles; a.training = d->training; a.eps = d->eps; a.momentum = d->momentum;
    if (!a.has_bn) {
      a.loss = d->loss; a.target = target; a.ignore = d->ignore_index;
      a.loss_part = reinterpret_cast<float*>(scratch + lo.loss_part);
    }
    ProfScope prof(KK_HEAD_FWD, st);
    static const bool narrow = [] { const char* e = getenv("OCGCN_HEAD_NARROW"); return !(e && e[0] == '0'); }();
    if (narrow && head_fwd_narrow_ok(a.Kin, a.Nout, a.has_bn, a.loss)) {     // 32-column CTAs: four times the CTAs of the wide kernel at width 128
      ST_TRY((ensure_dynamic_smem<head_fwd_narrow_kernel>(head_fwd_narrow_smem(kHeadMaxWidth))));
      CU_TRY(launch_chain(head_fwd_narrow_kernel, dim3(lo.grid, (a.Nout + HNC - 1) / HNC), dim3(HT), head_fwd_narrow_smem(a.Kin), st, a));
    } else {
      const size_t smem = head_fwd_smem(a.Kin, a.Nout);
      auto kern = a.Nout > 128 ? head_fwd_kernel<2> : head_fwd_kernel<1>;
      CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head_fwd_smem(kHeadMaxWidth, kHeadMaxWidth)));
      CU_TRY(launch_chain(kern, dim3(lo.grid), dim3(HT), smem, st, a));
    }
    LAUNCH_CHECK();
  }
  if (d->loss != OCGCN_LOSS_NONE) {
    ProfScope prof(KK_HEAD_LOSS, st);
    CU_TRY(launch_chain(head_loss_kernel, dim3(1), dim3(HT), 0, st, reinterpret_cast<const float*>(scratch + lo.loss_part), lo.ntiles, lo.T, d->B, d->loss, loss,
                                       reinterpret_cast<float*>(saved + lo.ncount)));
    LAUNCH_CHECK();
  }
  return OCGCN_OK;
}

int ocgcn_head_backward(const ocgcn_head_desc* d, const float* grad, const float* inp, const float* pred, const float* target,
                        const ocgcn_head_params* p, const ocgcn_head_grads* g, float* grad_inp, void* saved_v, void* scratch_v,
                        void* stream) {
  tl_launches = 0;
  tl_tile_parity = 0;
  HeadLayout lo;
  ST_TRY(head_layout(d, &lo));
  ST_TRY(head_validate(d, p));
  if (!g || !saved_v || !scratch_v) return OCGCN_EINVAL;
  for (int i = 0; i < lo.L; ++i) {
    if (!g->dW[i]) return OCGCN_EINVAL;
    if (i + 1 < lo.L && (!g->dgamma[i] || !g->dbeta[i])) return OCGCN_EINVAL;
  }
  if (d->need_dx && !grad_inp) return OCGCN_EINVAL;
  ST_TRY(check_device());
  ST_TRY(check_device_ptr(inp));
  ST_TRY(check_device_ptr(pred));
  if (d->loss == OCGCN_LOSS_NONE) ST_TRY(check_device_ptr(grad));
  else ST_TRY(check_device_ptr(target));
  if (((uintptr_t)saved_v & 255) || ((uintptr_t)scratch_v & 255)) return OCGCN_EALIGN;
  char* saved = static_cast<char*>(saved_v);
  char* scratch = static_cast<char*>(scratch_v);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int L = lo.L;
  ReduceJobs rj;
  memset(&rj, 0, sizeof(rj));
  for (int i = L - 1; i >= 0; --i) {
    HeadBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.B = d->B; a.ntiles = lo.ntiles; a.training = d->training;
    a.W = p->W[i]; a.Nout = d->width[i + 1]; a.Kin = d->width[i];
    a.has_bn = i + 1 < L;
    if (a.has_bn) {
      a.dY = reinterpret_cast<const float*>(scratch + lo.dY[i]);
      a.dpart = reinterpret_cast<const float*>(scratch + lo.dpart[i]);
      a.U = reinterpret_cast<const float*>(saved + lo.U[i]);
      a.stat = reinterpret_cast<const float*>(saved + lo.stat[i]);
      a.gamma = p->gamma[i]; a.dgamma = g->dgamma[i]; a.dbeta = g->dbeta[i];
    } else {
      a.pred = pred; a.act = d->act[i]; a.loss = d->loss; a.target = target; a.ignore = d->ignore_index;
      if (d->loss == OCGCN_LOSS_NONE) a.grad_pred = grad; else a.gloss = grad;
      a.ncount = reinterpret_cast<const float*>(saved + lo.ncount);
    }
    a.prev = i == 0 ? inp : reinterpret_cast<const float*>(saved + lo.U[i - 1]);
    a.prev_bn = i > 0;
    if (i > 0) {
      a.pstat = reinterpret_cast<const float*>(saved + lo.stat[i - 1]);
      a.pgamma = p->gamma[i - 1]; a.pbeta = p->beta[i - 1]; a.pact = d->act[i - 1];
      a.dprev = reinterpret_cast<float*>(scratch + lo.dY[i - 1]);
      a.dpart_out = reinterpret_cast<float*>(scratch + lo.dpart[i - 1]);
    } else {
      a.dprev = d->need_dx ? grad_inp : nullptr;
    }
    a.dWpart = reinterpret_cast<float*>(scratch + lo.dWpart[i]);
    a.dbpart = reinterpret_cast<float*>(scratch + lo.dbpart[i]);
    {
      ProfScope prof(KK_HEAD_BWD, st);
      static const bool narrow = [] { const char* e = getenv("OCGCN_HEAD_NARROW"); return !(e && e[0] == '0'); }();
      // 32-input-column CTAs (see head_bwd_narrow_kernel) for batches of at most 32 row tiles: each column block repeats the dU tile and
      // the fold of the per-tile partials, which costs more than it hides once the row tiles alone fill the GPU (C3: 36 vs 28 us)
      if (narrow && lo.ntiles <= 32 && head_bwd_narrow_ok(a.Kin, a.W)) {
        ST_TRY((ensure_dynamic_smem<head_bwd_narrow_kernel>(head_bwd_narrow_smem(kHeadMaxWidth))));
        CU_TRY(launch_chain(head_bwd_narrow_kernel, dim3(lo.grid, (a.Kin + HNC - 1) / HNC), dim3(HT), head_bwd_narrow_smem(a.Nout), st, a));
      } else {
        const size_t smem = head_bwd_smem(a.Kin, a.Nout);
        auto kern = a.Kin > 128 ? head_bwd_kernel<2> : head_bwd_kernel<1>;
        CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head_bwd_smem(kHeadMaxWidth, kHeadMaxWidth)));
        CU_TRY(launch_chain(kern, dim3(lo.grid), dim3(HT), smem, st, a));
      }
      LAUNCH_CHECK();
    }
    add_reduce(rj, a.dWpart, g->dW[i], lo.grid, a.Nout * a.Kin, 1, 0);
    add_reduce(rj, a.dbpart, g->db[i], lo.grid, a.Nout, 0, 0);
  }
  return run_reduce(rj, st);
}

}  // extern "C"
