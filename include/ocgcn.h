/*
 * ocgcn.h — C ABI of the B200-native GraphCNN hot path (libocgcn.so).
 *
 * Every entry point replaces a piece of the OpenChem reference (paths relative to the reference root):
 *
 *   ocgcn_encoder_forward   <- GraphCNNEncoder.forward          openchem/modules/encoders/gcn_encoder.py:38-53
 *                              (per layer GraphConvolution.forward openchem/layers/gcn.py:33-42, tanh :43,
 *                               neighbour max-pool :44-50, dense/tanh/sum/tanh readout :51-53)
 *   ocgcn_encoder_backward  <- autograd of the above (the reference has no explicit backward)
 *   ocgcn_layer_forward     <- GraphConvolution.forward          openchem/layers/gcn.py:33-42
 *   ocgcn_layer_backward    <- autograd of GraphConvolution.forward (dX, dW, db, dgamma, dbeta)
 *   ocgcn_query_workspace   <- (implicit) the temporaries autograd keeps alive in the reference
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer on the current CUDA device unless noted;
 *   - the caller owns every buffer (inputs, outputs, grads, running stats, workspaces); the library never
 *     allocates device memory and keeps no per-call state;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no device synchronisation inside;
 *   - thread-safe / re-entrant (DataParallel drives one replica per host thread, run.py:238);
 *   - return value: 0 = OK, otherwise an OCGCN_E* code (ocgcn_status_string gives text). No C++ exceptions
 *     cross the boundary. There is NO CPU fallback: host pointers are the caller's error.
 */
#ifndef OCGCN_H_
#define OCGCN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCGCN_MAX_LAYERS 16
#define OCGCN_MAX_ATOMS 256 /* arg-max indices are stored as u8 */

/* precision classes (encoder_params['precision']) */
#define OCGCN_FP32 0 /* exact fp32 FFMA products, fp32 accumulate               (1e-5 class) */
#define OCGCN_TF32 1 /* tensor-core products, split operands in the forward     (1e-2 class) */
#define OCGCN_BF16 2 /* tensor-core products, split operands in the forward     (1e-2 class) */

/* status codes */
#define OCGCN_OK 0
#define OCGCN_EINVAL 1      /* bad shape / null pointer / bad enum */
#define OCGCN_EUNSUPPORTED 2 /* size or mode outside what the kernels implement */
#define OCGCN_EALIGN 3      /* misaligned pointer */
#define OCGCN_EARCH 4       /* not an sm_100 device */
#define OCGCN_ECUDA 5       /* a CUDA call failed (see ocgcn_last_cuda_error) */

typedef struct ocgcn_desc {
  int32_t B;                       /* molecules in the batch                                  */
  int32_t N;                       /* padded atoms per molecule (tensor dim 1), <= 256        */
  int32_t L;                       /* graph-conv layers (1 for the layer_* entry points)      */
  int32_t E;                       /* encoder_dim (dense/readout width); unused by layer_*    */
  int32_t F[OCGCN_MAX_LAYERS + 1]; /* F[0] = input_size, F[l] = hidden_size[l-1]              */
  int32_t mode;                    /* OCGCN_FP32 / OCGCN_TF32 / OCGCN_BF16                     */
  int32_t training;                /* 1: batch statistics + running-stat update; 0: running   */
  int32_t has_bias;                /* GraphConvolution(bias=...)                              */
  int32_t need_dx;                 /* layer_backward: also produce dX = A^T (dZ W^T)          */
  float eps;                       /* BatchNorm1d eps (1e-5)                                  */
  float momentum;                  /* BatchNorm1d momentum (0.1)                              */
  int64_t x_stride[3];             /* element strides of x  (B,N,F[0])                        */
  int64_t adj_stride[3];           /* element strides of adj (B,N,N)                          */
} ocgcn_desc;

/* Parameter / buffer pointers, indexed by layer. Names follow the reference state_dict:
 * graph_convolutions.{l}.weight (F[l],F[l+1]) [row-major, NOT nn.Linear layout, gcn.py:19],
 * .bias (F[l+1]), .bn.weight, .bn.bias, .bn.running_mean, .bn.running_var, .bn.num_batches_tracked (int64 scalar),
 * dense.weight (E,F[L]) [nn.Linear layout], dense.bias (E). */
typedef struct ocgcn_params {
  const float* W[OCGCN_MAX_LAYERS];
  const float* b[OCGCN_MAX_LAYERS]; /* NULL entries when has_bias == 0 */
  const float* gamma[OCGCN_MAX_LAYERS];
  const float* beta[OCGCN_MAX_LAYERS];
  float* running_mean[OCGCN_MAX_LAYERS];         /* updated in place when training */
  float* running_var[OCGCN_MAX_LAYERS];          /* updated in place when training */
  int64_t* num_batches_tracked[OCGCN_MAX_LAYERS]; /* += 1 when training (may be NULL) */
  const float* Wd;
  const float* bd;
} ocgcn_params;

typedef struct ocgcn_grads {
  float* dW[OCGCN_MAX_LAYERS];
  float* db[OCGCN_MAX_LAYERS]; /* NULL entries when has_bias == 0 */
  float* dgamma[OCGCN_MAX_LAYERS];
  float* dbeta[OCGCN_MAX_LAYERS];
  float* dWd;
  float* dbd;
} ocgcn_grads;

/* Bytes of the two caller-provided buffers:
 *   saved_bytes   — written by *_forward, read by *_backward (keep it alive in the autograd ctx);
 *   scratch_bytes — transient, needed only for the duration of one *_forward or *_backward call.
 * Valid for both encoder_* (desc->L layers + readout) and layer_* (desc->L == 1) calls. 256-byte aligned. */
int ocgcn_query_workspace(const ocgcn_desc* desc, size_t* saved_bytes, size_t* scratch_bytes);

/* Where a tensor lives inside the `saved` buffer (for tests / debugging; the layout is otherwise private):
 * Z = pre-BatchNorm activations (B*N, F[layer+1]) fp32; ARGMAX = pool arg-max j* (B*N, F[layer+1]) u8;
 * S = adj@H_{layer} (B*N, F[layer]) fp32; STATS = mean, invstd, scale, shift (4, F[layer+1]) fp32;
 * ROWMASK = adjacency bit rows (B*N, ceil(N/64)) u64; MOLFLAG = per-molecule "adjacency is exactly 0/1" (B) u32. */
#define OCGCN_SAVED_Z 0
#define OCGCN_SAVED_ARGMAX 1
#define OCGCN_SAVED_S 2
#define OCGCN_SAVED_STATS 3
#define OCGCN_SAVED_ROWMASK 4
#define OCGCN_SAVED_MOLFLAG 5
int ocgcn_saved_tensor(const ocgcn_desc* desc, int what, int layer, size_t* offset, size_t* bytes);

/* out (B,E) contiguous fp32. x (B,N,F[0]) and adj (B,N,N) fp32 with the strides in desc. */
int ocgcn_encoder_forward(const ocgcn_desc* desc, const float* x, const float* adj, const ocgcn_params* params,
                          void* saved, void* scratch, float* out, void* stream);

/* grad_out (B,E) contiguous. Writes (does not accumulate) every gradient in `grads`. */
int ocgcn_encoder_backward(const ocgcn_desc* desc, const float* grad_out, const float* x, const float* adj,
                           const ocgcn_params* params, const ocgcn_grads* grads, void* saved, void* scratch,
                           void* stream);

/* Standalone GraphConvolution: y (B,N,F[1]) contiguous = BN((adj x) W + b). Uses layer index 0 of params. */
int ocgcn_layer_forward(const ocgcn_desc* desc, const float* x, const float* adj, const ocgcn_params* params,
                        void* saved, void* scratch, float* y, void* stream);

/* grad_y (B,N,F[1]) contiguous. grad_x (B,N,F[0]) contiguous, written when desc->need_dx. Layer index 0 of grads. */
int ocgcn_layer_backward(const ocgcn_desc* desc, const float* grad_y, const float* x, const float* adj,
                         const ocgcn_params* params, const ocgcn_grads* grads, float* grad_x, void* saved,
                         void* scratch, void* stream);

/* Single-kernel EVAL forward (desc->training == 0) of the whole encoder, gcn_encoder.py:38-53 with BatchNorm on its running
 * statistics: every layer of a molecule tile runs inside one persistent CTA and the activations stay on the SM between layers
 * (csrc/ocgcn_fused.cuh). Nothing is saved for a backward pass: a caller that later needs gradients re-runs ocgcn_encoder_forward
 * (same results, bit for bit) and then ocgcn_encoder_backward. Served: tensor-core classes, N <= 128, every width and E <= 128,
 * L <= 8; anything else returns OCGCN_EUNSUPPORTED from both functions and the caller uses ocgcn_encoder_forward.
 * `workspace` (ocgcn_query_eval_workspace bytes, 256-byte aligned) holds the packed adjacency and the weight images for the call. */
int ocgcn_query_eval_workspace(const ocgcn_desc* desc, size_t* bytes);
int ocgcn_encoder_forward_eval(const ocgcn_desc* desc, const float* x, const float* adj, const ocgcn_params* params, void* workspace,
                               float* out, void* stream);

/* tanh + neighbour max-pool as a stand-alone differentiable op (replaces torch.tanh + x.repeat/view/expand/max(dim=2) of
 * openchem/modules/encoders/edge_attention_encoder.py:54-61 and gcn_encoder.py:43-50 for callers that chain GraphConvolution layers
 * themselves): pooled[b,i,c] = max_j adj[b,i,j] * tanh(y[b,j,c]), first index on ties; argmax (B,N,D) uint8 is the saved routing.
 * backward: grad_y = (scatter of grad_pooled through argmax, times adj) * (1 - tanh(y)^2). y / pooled / grads contiguous (B,N,D);
 * adj (B,N,N) with element strides adj_stride[3]; N <= OCGCN_MAX_ATOMS. Exact fp32, fixed summation order. */
int ocgcn_pool_forward(int B, int N, int D, const float* y, const float* adj, const int64_t* adj_stride, float* pooled,
                       uint8_t* argmax, void* stream);
int ocgcn_pool_backward(int B, int N, int D, const float* y, const float* adj, const int64_t* adj_stride, const uint8_t* argmax,
                        const float* grad_pooled, float* grad_y, void* stream);

/* Compact batch format (replaces the dense fp32 upload of Graph2Label.cast_inputs, openchem/models/Graph2Label.py:41-57,
 * for 0/1 adjacency and 0/1 node features — what utils/graph.py:57-63,80-84 produces). Bit (c & 31) of word
 * bits[(mol*N + row)*ceil(C/32) + (c >> 5)] is set iff dense[mol][row][c] == 1.0f; C = N for adj_bits, F0 for x_bits.
 *
 * ocgcn_expand_compact writes the contiguous dense fp32 x (B,N,F0) and adj (B,N,N) the entry points above take (bit-exact).
 * `index` (device, B int32, or NULL) gathers molecule index[b] of a compact data set resident in device memory into
 * batch slot b; the data set holds n_src molecules. An index outside [0, n_src) is never dereferenced: that molecule is written
 * as zeros and *index_oob (device int32, zeroed by the call; required with `index`) is set to 1. Either output (with its bits)
 * may be NULL to skip it.
 * ocgcn_pack_compact is the inverse for strided dense inputs (element strides, molecule/row/column); *nonbinary (device
 * int32, zeroed by the call) is set to 1 if any entry is neither 0 nor 1 — such batches must use the dense entry points. */
int ocgcn_expand_compact(int B, int N, int F0, const uint32_t* x_bits, const uint32_t* adj_bits, const int32_t* index, int n_src,
                         int32_t* index_oob, float* x, float* adj, void* stream);
int ocgcn_pack_compact(int B, int N, int F0, const float* x, const int64_t* x_stride, const float* adj,
                       const int64_t* adj_stride, uint32_t* x_bits, uint32_t* adj_bits, int32_t* nonbinary, void* stream);

/* The encoder entry points fed with the bit words themselves (same layout as above; desc->x_stride / adj_stride are ignored): the
 * neighbour records are built straight from adj_bits and the first layer reads x_bits, so the dense fp32 x / adj (8*N*N bytes per
 * molecule written by ocgcn_expand_compact and read again by the forward) never exist. Results are bit-identical to
 * ocgcn_expand_compact followed by the dense entry point (GraphCNNEncoder.forward on the same batch, gcn_encoder.py:38-53).
 * The backward of a *_forward_compact call is ocgcn_encoder_backward_compact (same `saved`); it needs neither x nor adj. */
int ocgcn_encoder_forward_compact(const ocgcn_desc* desc, const uint32_t* x_bits, const uint32_t* adj_bits,
                                  const ocgcn_params* params, void* saved, void* scratch, float* out, void* stream);
int ocgcn_encoder_forward_eval_compact(const ocgcn_desc* desc, const uint32_t* x_bits, const uint32_t* adj_bits,
                                       const ocgcn_params* params, void* workspace, float* out, void* stream);
int ocgcn_encoder_backward_compact(const ocgcn_desc* desc, const float* grad_out, const ocgcn_params* params,
                                   const ocgcn_grads* grads, void* saved, void* scratch, void* stream);

/* ---- MLP head + loss (Graph2Label.forward's `self.MLP(output)` and the criterion `fit` applies to it) -------------------
 * ocgcn_head_forward   <- OpenChemMLP.forward                      openchem/modules/mlp/openchem_mlp.py:51-61
 *                         (+ nn.MSELoss / MultitaskLoss.forward    openchem/criterion/multitask_loss.py:40-49 when loss != NONE)
 * ocgcn_head_backward  <- autograd of the above
 * Layer i: Linear(width[i] -> width[i+1]); every layer but the last is followed by BatchNorm1d over the B rows and act[i],
 * the last by act[n_layers-1] only. Exact fp32 arithmetic. Dropout must be inactive (p = 0 or eval mode). */
#define OCGCN_HEAD_MAX_LAYERS 4
#define OCGCN_HEAD_MAX_WIDTH 256
#define OCGCN_ACT_IDENTITY 0
#define OCGCN_ACT_RELU 1
#define OCGCN_ACT_SIGMOID 2
#define OCGCN_ACT_TANH 3
#define OCGCN_LOSS_NONE 0      /* forward returns predictions only; backward takes grad_pred (B,T)                     */
#define OCGCN_LOSS_MSE 1       /* nn.MSELoss(): mean over B*T                                                         */
#define OCGCN_LOSS_MULTITASK 2 /* MultitaskLoss(ignore_index, n_tasks = T): masked BCE, per-task mean, mean over tasks */

typedef struct ocgcn_head_desc {
  int32_t B;
  int32_t n_layers;                          /* 1..OCGCN_HEAD_MAX_LAYERS                                  */
  int32_t width[OCGCN_HEAD_MAX_LAYERS + 1];  /* input_size, hidden_size[0..n_layers-1] (last = T)         */
  int32_t act[OCGCN_HEAD_MAX_LAYERS];        /* OCGCN_ACT_*                                               */
  int32_t training;                          /* BatchNorm: batch statistics + running-stat update         */
  int32_t loss;                              /* OCGCN_LOSS_*                                              */
  int32_t need_dx;                           /* backward: also produce grad_inp (B, width[0])             */
  float eps, momentum;                       /* BatchNorm1d                                               */
  float ignore_index;                        /* MultitaskLoss                                             */
} ocgcn_head_desc;

/* state_dict names: layers.{i}.weight (width[i+1], width[i]) [nn.Linear layout], layers.{i}.bias, bn.{i}.weight/.bias/
 * .running_mean/.running_var/.num_batches_tracked for i < n_layers-1 */
typedef struct ocgcn_head_params {
  const float* W[OCGCN_HEAD_MAX_LAYERS];
  const float* b[OCGCN_HEAD_MAX_LAYERS];
  const float* gamma[OCGCN_HEAD_MAX_LAYERS];
  const float* beta[OCGCN_HEAD_MAX_LAYERS];
  float* running_mean[OCGCN_HEAD_MAX_LAYERS];
  float* running_var[OCGCN_HEAD_MAX_LAYERS];
  int64_t* num_batches_tracked[OCGCN_HEAD_MAX_LAYERS];
} ocgcn_head_params;

typedef struct ocgcn_head_grads {
  float* dW[OCGCN_HEAD_MAX_LAYERS];
  float* db[OCGCN_HEAD_MAX_LAYERS];
  float* dgamma[OCGCN_HEAD_MAX_LAYERS];
  float* dbeta[OCGCN_HEAD_MAX_LAYERS];
} ocgcn_head_grads;

int ocgcn_head_query_workspace(const ocgcn_head_desc* desc, size_t* saved_bytes, size_t* scratch_bytes);
/* inp (B,width[0]) contiguous; pred (B,T) written; target (B,T) and loss (device scalar) used when desc->loss != NONE. */
int ocgcn_head_forward(const ocgcn_head_desc* desc, const float* inp, const ocgcn_head_params* params, const float* target,
                       void* saved, void* scratch, float* pred, float* loss, void* stream);
/* grad: desc->loss == NONE: grad_pred (B,T); otherwise a device scalar dLoss (NULL = 1). pred = what forward wrote.
 * Writes every gradient in `grads` and, when desc->need_dx, grad_inp (B,width[0]). */
int ocgcn_head_backward(const ocgcn_head_desc* desc, const float* grad, const float* inp, const float* pred, const float* target,
                        const ocgcn_head_params* params, const ocgcn_head_grads* grads, float* grad_inp, void* saved,
                        void* scratch, void* stream);

/* Opt-in, process-wide per-kernel timing (used by bench.py's roofline leg, never by training):
 * while enabled every kernel launch is bracketed by CUDA events recorded on the launching stream.
 * ocgcn_profile_collect synchronises those events, sums milliseconds and launch counts per kernel kind
 * (ocgcn_profile_kinds() kinds, names via ocgcn_profile_kind_name) and clears the record. */
int ocgcn_profile_enable(int on);
int ocgcn_profile_kinds(void);
const char* ocgcn_profile_kind_name(int kind);
int ocgcn_profile_collect(float* ms_by_kind, int* launches_by_kind, int n_kinds);

/* Number of kernel launches the last call on this host thread enqueued (bench.py's gpu_launches claim). */
int ocgcn_last_launch_count(void);
/* cudaError_t of the last failing CUDA call on this host thread (0 if none). */
int ocgcn_last_cuda_error(void);
const char* ocgcn_status_string(int status);
/* ABI version; bumped whenever a struct above changes. */
/* Bumped whenever a struct layout or an entry-point signature changes (2: ocgcn_expand_compact gained n_src / index_oob; 3: ocgcn_pool_forward / _backward; 4: ocgcn_encoder_forward_eval; 5: ocgcn_encoder_*_compact). Bindings must refuse a library that reports another value. */
#define OCGCN_ABI_VERSION 5
int ocgcn_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* OCGCN_H_ */
