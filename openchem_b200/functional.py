"""autograd.Function wrappers that marshal torch tensors into the C ABI (include/ocgcn.h).

PyTorch is plumbing here: device memory (workspaces live in torch tensors owned by the autograd ctx),
the current stream, and the autograd graph. All arithmetic happens in libocgcn.so.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib

# launches enqueued by this process through the C ABI (bench.py reports it as gpu_launches)
launch_counter = {"n": 0}


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("openchem_b200: %s must be a CUDA tensor (this path has no CPU fallback)" % name)


def _f32(t):
    return t if t.dtype == torch.float32 else t.to(torch.float32)


def _check_state(ref, named):
    """Every parameter / BatchNorm buffer handed to the kernels by raw pointer must live on the inputs' device, in the dtype the
    kernels read (float32; num_batches_tracked int64) and contiguous - the same class of error the reference raises from ATen
    when a module was left on the CPU, on another GPU, or cast with .double()/.half(). Without this the kernels would
    dereference host or wrong-typed memory."""
    for name, t, dtype in named:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.device != ref.device:
            raise RuntimeError("openchem_b200: %s is on %s but the input is on %s (move the module with .to(device); this path has "
                               "no CPU fallback)" % (name, getattr(t, "device", type(t).__name__), ref.device))
        if t.dtype != dtype:
            raise RuntimeError("openchem_b200: %s has dtype %s, the kernels need %s" % (name, t.dtype, dtype))
        if not t.is_contiguous():
            raise RuntimeError("openchem_b200: %s must be contiguous" % name)


def _layer_state(spec, Ws, bs, gammas, betas, Wd=None, bd=None):
    named = []
    for l in range(len(Ws)):
        named += [("weight[%d]" % l, Ws[l], torch.float32), ("bias[%d]" % l, bs[l], torch.float32),
                  ("bn.weight[%d]" % l, gammas[l], torch.float32), ("bn.bias[%d]" % l, betas[l], torch.float32),
                  ("bn.running_mean[%d]" % l, spec.running_mean[l], torch.float32),
                  ("bn.running_var[%d]" % l, spec.running_var[l], torch.float32),
                  ("bn.num_batches_tracked[%d]" % l, spec.nbt[l], torch.int64)]
    named += [("dense.weight", Wd, torch.float32), ("dense.bias", bd, torch.float32)]
    return named


class LayerSpec:
    """Non-tensor arguments + in-place updated buffers of one call (kept out of autograd's sight)."""

    def __init__(self, widths, encoder_dim, mode, training, has_bias, eps, momentum, running_mean, running_var, nbt):
        self.widths = list(widths)          # [F0, F1, ..., FL]
        self.encoder_dim = int(encoder_dim)  # 0 => standalone layer call
        self.mode = mode
        self.training = bool(training)
        self.has_bias = bool(has_bias)
        self.eps = float(eps)
        self.momentum = float(momentum)
        self.running_mean = running_mean
        self.running_var = running_var
        self.nbt = nbt


def is_compact_pair(x, adj):
    """(x, adj) are the int32 bit words of an openchem_b200.compact.CompactGraphBatch rather than dense fp32 tensors."""
    return x.dtype == torch.int32 and adj.dtype == torch.int32


def _make_desc(spec, x, adj, need_dx=False):
    d = _lib.Desc()
    if is_compact_pair(x, adj):      # bit words (B,N,ceil(F0/32)) and (B,N,ceil(N/32)); the strides in the descriptor are not used
        B, N, XW = x.shape
        F0 = spec.widths[0]
        if tuple(adj.shape) != (B, N, (N + 31) // 32) or XW != (F0 + 31) // 32 or not x.is_contiguous() or not adj.is_contiguous():
            raise RuntimeError("openchem_b200: compact batch needs contiguous x_bits (B,N,ceil(F0/32)) and adj_bits (B,N,ceil(N/32)) "
                               "with F0 = %d; got %s and %s" % (F0, tuple(x.shape), tuple(adj.shape)))
    else:
        B, N, F0 = x.shape
        if adj.dim() != 3 or adj.shape[0] != B or adj.shape[1] != N or adj.shape[2] != N:
            raise RuntimeError("openchem_b200: adj must be (B,N,N) matching x (B,N,F); got %s and %s" % (tuple(adj.shape), tuple(x.shape)))
    if F0 != spec.widths[0]:
        raise RuntimeError("openchem_b200: x has %d features, layer expects %d" % (F0, spec.widths[0]))
    d.B, d.N, d.L, d.E = B, N, len(spec.widths) - 1, spec.encoder_dim
    for i, f in enumerate(spec.widths):
        d.F[i] = f
    d.mode = _lib.MODES[spec.mode]
    d.training = int(spec.training)
    d.has_bias = int(spec.has_bias)
    d.need_dx = int(need_dx)
    d.eps, d.momentum = spec.eps, spec.momentum
    for i in range(3):
        d.x_stride[i] = x.stride(i)
        d.adj_stride[i] = adj.stride(i)
    if is_compact_pair(x, adj):
        for i, (sx, sa) in enumerate(((N * F0, N * N), (F0, N), (1, 1))):
            d.x_stride[i], d.adj_stride[i] = sx, sa
    return d


def _make_params(spec, Ws, bs, gammas, betas, Wd=None, bd=None):
    p = _lib.Params()
    for l in range(len(Ws)):
        p.W[l] = Ws[l].data_ptr()
        p.b[l] = bs[l].data_ptr() if bs[l] is not None else None
        p.gamma[l] = gammas[l].data_ptr()
        p.beta[l] = betas[l].data_ptr()
        p.running_mean[l] = spec.running_mean[l].data_ptr()
        p.running_var[l] = spec.running_var[l].data_ptr()
        p.num_batches_tracked[l] = spec.nbt[l].data_ptr() if spec.nbt[l] is not None else None
    if Wd is not None:
        p.Wd, p.bd = Wd.data_ptr(), bd.data_ptr()
    return p


def _workspaces(desc, device):
    lib = _lib.load()
    sb, cb = ctypes.c_size_t(0), ctypes.c_size_t(0)
    _lib.check(lib.ocgcn_query_workspace(ctypes.byref(desc), ctypes.byref(sb), ctypes.byref(cb)))
    return sb.value, cb.value


def _alloc(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _count(lib):
    launch_counter["n"] += lib.ocgcn_last_launch_count()


# Gradient sinks (openchem_b200.step.TrainStep): inside a sink scope the backward kernels write each parameter gradient straight into a
# caller-registered buffer (a view of the step's flat gradient) and return that view, so a `torch.autograd.grad` call needs no
# per-parameter accumulate / copy kernel (28 tiny launches per step at the GraphCNN shapes). The decision is taken at FORWARD time in
# the calling thread; outside a scope (plain `loss.backward()`, run.py's train_step) fresh tensors are returned as autograd expects.
from torch.utils.weak import WeakIdKeyDictionary  # noqa: E402  (keys compared by identity: Tensor.__eq__ is elementwise)

grad_sinks = WeakIdKeyDictionary()            # nn.Parameter -> destination tensor (same shape, contiguous, same device)
sink_scope = {"on": False}


def _sinks_for(params):
    if not sink_scope["on"]:
        return None
    return [None if t is None else grad_sinks.get(t) for t in params]


def _grad_buffer(like, sinks, i):
    if sinks is not None and sinks[i] is not None:
        return sinks[i]
    return torch.empty_like(like)


# single-kernel eval forward (csrc/ocgcn_fused.cuh): on by default at every batch size (measured on a B200, bf16 class, eval forward
# of the encoder: C2 60 vs 131 us layer by layer, C1 107 vs 164 us, C4 312 vs 480 us, C3 654 vs 870 us; profiles/r2_fused_eval.md);
# OCB200_FUSED_EVAL=0 or `max_tiles` restrict it
fused_eval = {"on": os.environ.get("OCB200_FUSED_EVAL", "1") != "0", "max_tiles": int(os.environ.get("OCB200_FUSED_EVAL_MAX_TILES", str(1 << 30)))}


def fused_eval_tiles(desc):
    return (desc.B + (128 // desc.N) - 1) // (128 // desc.N) if 0 < desc.N <= 128 else 1 << 30


SAVED_KINDS = {"Z": (0, torch.float32), "argmax": (1, torch.uint8), "S": (2, torch.float32), "stats": (3, torch.float32),
               "rowmask": (4, torch.int64), "molflag": (5, torch.int32)}
compact_direct = {"on": True}    # GraphCNNEncoder.forward(CompactGraphBatch): feed the bit words to the kernels (off: expand to dense first)
debug_keep = {"on": False, "saved": None, "desc": None}     # tests only: keep the last forward's saved buffer


def saved_tensor(kind, layer=0):
    """(tests) view of one tensor inside the last forward's saved workspace; needs debug_keep['on'] = True."""
    lib = _lib.load()
    what, dtype = SAVED_KINDS[kind]
    off, nb = ctypes.c_size_t(0), ctypes.c_size_t(0)
    _lib.check(lib.ocgcn_saved_tensor(ctypes.byref(debug_keep["desc"]), what, layer, ctypes.byref(off), ctypes.byref(nb)))
    return debug_keep["saved"][off.value: off.value + nb.value].view(dtype)


class EncoderFn(torch.autograd.Function):
    """GraphCNNEncoder.forward (gcn_encoder.py:38-53) as ONE autograd node. inputs: x, adj, spec, then per layer
    W, b, gamma, beta (b may be None) and finally Wd, bd. (x, adj) may also be the int32 bit words of a compact batch
    (compact.CompactGraphBatch.x_bits / .adj_bits): the ocgcn_encoder_*_compact entry points then read the words directly."""

    @staticmethod
    def forward(ctx, x, adj, spec, *tensors):
        lib = _lib.load()
        if (x.dtype == torch.int32) != (adj.dtype == torch.int32):
            raise RuntimeError("openchem_b200: x and adj must BOTH be int32 bit words of a compact batch or both dense tensors "
                               "(got %s and %s)" % (x.dtype, adj.dtype))
        _require_cuda(x, "x")
        _require_cuda(adj, "adj")
        if x.requires_grad or adj.requires_grad:
            raise RuntimeError("openchem_b200: the fused encoder treats x and adj as data (no gradient); "
                               "use GraphConvolution layers if you need d/dx")
        L = len(spec.widths) - 1
        compact = is_compact_pair(x, adj)
        if not compact:
            x, adj = _f32(x), _f32(adj)
        ctx.sinks = _sinks_for(tensors)
        tensors = [None if t is None else t.detach().contiguous() for t in tensors]
        Ws, bs, gammas, betas = tensors[0:4 * L:4], tensors[1:4 * L:4], tensors[2:4 * L:4], tensors[3:4 * L:4]
        Wd, bd = tensors[4 * L], tensors[4 * L + 1]
        if adj.device != x.device:
            raise RuntimeError("openchem_b200: x is on %s but adj is on %s" % (x.device, adj.device))
        _check_state(x, _layer_state(spec, Ws, bs, gammas, betas, Wd, bd))
        with torch.cuda.device(x.device):
            desc = _make_desc(spec, x, adj)
            params = _make_params(spec, Ws, bs, gammas, betas, Wd, bd)
            out = torch.empty((x.shape[0], spec.encoder_dim), dtype=torch.float32, device=x.device)
            saved, cbytes = None, 0
            ebytes = ctypes.c_size_t(0)
            # eval mode: the whole stack in ONE kernel, activations stay on the SM between layers (csrc/ocgcn_fused.cuh). Nothing is
            # saved; if a backward pass is requested later (the reference's evaluate() runs without no_grad) the saved state is
            # rebuilt then by the layer-by-layer forward, which gives bit-identical results.
            if (not spec.training and fused_eval["on"] and not debug_keep["on"]
                    and lib.ocgcn_query_eval_workspace(ctypes.byref(desc), ctypes.byref(ebytes)) == 0
                    and fused_eval_tiles(desc) <= fused_eval["max_tiles"]):
                ws = _alloc(ebytes.value, x.device)
                fwd_eval = lib.ocgcn_encoder_forward_eval_compact if compact else lib.ocgcn_encoder_forward_eval
                _lib.check(fwd_eval(ctypes.byref(desc), x.data_ptr(), adj.data_ptr(), ctypes.byref(params),
                                    ws.data_ptr(), out.data_ptr(), _stream_ptr(x.device)))
                _count(lib)
            else:
                sbytes, cbytes = _workspaces(desc, x.device)
                saved = _alloc(sbytes, x.device)
                scratch = _alloc(cbytes, x.device)
                fwd = lib.ocgcn_encoder_forward_compact if compact else lib.ocgcn_encoder_forward
                _lib.check(fwd(ctypes.byref(desc), x.data_ptr(), adj.data_ptr(), ctypes.byref(params),
                               saved.data_ptr(), scratch.data_ptr(), out.data_ptr(), _stream_ptr(x.device)))
                _count(lib)
        if debug_keep["on"]:
            debug_keep["saved"], debug_keep["desc"] = saved, desc
        ctx.spec = spec
        ctx.cbytes = cbytes
        ctx.lazy = saved is None
        ctx.save_for_backward(x, adj, *([] if saved is None else [saved]), *[t for t in tensors if t is not None])
        ctx.none_mask = [t is None for t in tensors]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        spec = ctx.spec
        L = len(spec.widths) - 1
        if ctx.lazy:
            x, adj, *rest = ctx.saved_tensors
            saved = None
        else:
            x, adj, saved, *rest = ctx.saved_tensors
        it = iter(rest)
        tensors = [None if isnone else next(it) for isnone in ctx.none_mask]
        Ws, bs, gammas, betas = tensors[0:4 * L:4], tensors[1:4 * L:4], tensors[2:4 * L:4], tensors[3:4 * L:4]
        Wd, bd = tensors[4 * L], tensors[4 * L + 1]
        grad_out = _f32(grad_out).contiguous()
        with torch.cuda.device(x.device):
            desc = _make_desc(spec, x, adj)
            params = _make_params(spec, Ws, bs, gammas, betas, Wd, bd)
            if saved is None:     # the forward was the single-kernel eval forward: rebuild the saved state now (eval mode: no side effects)
                sbytes, ctx.cbytes = _workspaces(desc, x.device)
                saved = _alloc(sbytes, x.device)
                tmp = torch.empty((x.shape[0], spec.encoder_dim), dtype=torch.float32, device=x.device)
                fwd = lib.ocgcn_encoder_forward_compact if is_compact_pair(x, adj) else lib.ocgcn_encoder_forward
                _lib.check(fwd(ctypes.byref(desc), x.data_ptr(), adj.data_ptr(), ctypes.byref(params),
                               saved.data_ptr(), _alloc(ctx.cbytes, x.device).data_ptr(), tmp.data_ptr(), _stream_ptr(x.device)))
                _count(lib)
            grads = _lib.Grads()
            outs = []
            sk = ctx.sinks
            for l in range(L):
                gW, gg, gb_ = _grad_buffer(Ws[l], sk, 4 * l), _grad_buffer(gammas[l], sk, 4 * l + 2), _grad_buffer(betas[l], sk, 4 * l + 3)
                gbias = _grad_buffer(bs[l], sk, 4 * l + 1) if bs[l] is not None else None
                grads.dW[l], grads.dgamma[l], grads.dbeta[l] = gW.data_ptr(), gg.data_ptr(), gb_.data_ptr()
                grads.db[l] = gbias.data_ptr() if gbias is not None else None
                outs += [gW, gbias, gg, gb_]
            gWd, gbd = _grad_buffer(Wd, sk, 4 * L), _grad_buffer(bd, sk, 4 * L + 1)
            grads.dWd, grads.dbd = gWd.data_ptr(), gbd.data_ptr()
            scratch = _alloc(ctx.cbytes, x.device)
            if is_compact_pair(x, adj):
                _lib.check(lib.ocgcn_encoder_backward_compact(ctypes.byref(desc), grad_out.data_ptr(), ctypes.byref(params),
                                                              ctypes.byref(grads), saved.data_ptr(), scratch.data_ptr(),
                                                              _stream_ptr(x.device)))
            else:
                _lib.check(lib.ocgcn_encoder_backward(ctypes.byref(desc), grad_out.data_ptr(), x.data_ptr(), adj.data_ptr(),
                                                      ctypes.byref(params), ctypes.byref(grads), saved.data_ptr(), scratch.data_ptr(),
                                                      _stream_ptr(x.device)))
            _count(lib)
        return (None, None, None, *outs, gWd, gbd)


class LayerFn(torch.autograd.Function):
    """GraphConvolution.forward (gcn.py:33-42) as one autograd node: inputs x, adj, spec, W, b, gamma, beta."""

    @staticmethod
    def forward(ctx, x, adj, spec, W, b, gamma, beta):
        lib = _lib.load()
        _require_cuda(x, "x")
        _require_cuda(adj, "adj")
        if adj.requires_grad:
            raise RuntimeError("openchem_b200: gradients w.r.t. the adjacency are not supported (no reference caller needs them)")
        need_dx = bool(x.requires_grad)
        x, adj = _f32(x.detach()), _f32(adj)
        W, gamma, beta = W.detach().contiguous(), gamma.detach().contiguous(), beta.detach().contiguous()
        b = None if b is None else b.detach().contiguous()
        if adj.device != x.device:
            raise RuntimeError("openchem_b200: x is on %s but adj is on %s" % (x.device, adj.device))
        _check_state(x, _layer_state(spec, [W], [b], [gamma], [beta]))
        with torch.cuda.device(x.device):
            desc = _make_desc(spec, x, adj, need_dx)
            params = _make_params(spec, [W], [b], [gamma], [beta])
            sbytes, cbytes = _workspaces(desc, x.device)
            saved = _alloc(sbytes, x.device)
            scratch = _alloc(cbytes, x.device)
            y = torch.empty((x.shape[0], x.shape[1], spec.widths[1]), dtype=torch.float32, device=x.device)
            _lib.check(lib.ocgcn_layer_forward(ctypes.byref(desc), x.data_ptr(), adj.data_ptr(), ctypes.byref(params),
                                               saved.data_ptr(), scratch.data_ptr(), y.data_ptr(), _stream_ptr(x.device)))
            _count(lib)
        ctx.spec, ctx.cbytes, ctx.need_dx, ctx.has_b = spec, cbytes, need_dx, b is not None
        ctx.save_for_backward(x, adj, saved, W, gamma, beta, *([b] if b is not None else []))
        return y

    @staticmethod
    def backward(ctx, grad_y):
        lib = _lib.load()
        spec = ctx.spec
        x, adj, saved, W, gamma, beta, *rest = ctx.saved_tensors
        b = rest[0] if ctx.has_b else None
        grad_y = _f32(grad_y).contiguous()
        with torch.cuda.device(x.device):
            desc = _make_desc(spec, x, adj, ctx.need_dx)
            params = _make_params(spec, [W], [b], [gamma], [beta])
            grads = _lib.Grads()
            gW, gg, gb_ = torch.empty_like(W), torch.empty_like(gamma), torch.empty_like(beta)
            gbias = torch.empty_like(b) if b is not None else None
            grads.dW[0], grads.dgamma[0], grads.dbeta[0] = gW.data_ptr(), gg.data_ptr(), gb_.data_ptr()
            grads.db[0] = gbias.data_ptr() if gbias is not None else None
            gx = torch.empty((x.shape[0], x.shape[1], spec.widths[0]), dtype=torch.float32, device=x.device) if ctx.need_dx else None
            scratch = _alloc(ctx.cbytes, x.device)
            _lib.check(lib.ocgcn_layer_backward(ctypes.byref(desc), grad_y.data_ptr(), x.data_ptr(), adj.data_ptr(),
                                                ctypes.byref(params), ctypes.byref(grads), gx.data_ptr() if gx is not None else None,
                                                saved.data_ptr(), scratch.data_ptr(), _stream_ptr(x.device)))
            _count(lib)
        return gx, None, None, gW, gbias, gg, gb_


# ---------------------------------------------------------------------------------------------------------------------
# MLP head (+ criterion) — ocgcn_head_forward / ocgcn_head_backward
# ---------------------------------------------------------------------------------------------------------------------
class HeadSpec:
    """Non-tensor arguments + in-place updated BatchNorm buffers of one head call."""

    def __init__(self, widths, acts, training, eps, momentum, running_mean, running_var, nbt, loss="none", ignore_index=0.0):
        self.widths, self.acts = [int(w) for w in widths], list(acts)
        self.training, self.eps, self.momentum = bool(training), float(eps), float(momentum)
        self.running_mean, self.running_var, self.nbt = running_mean, running_var, nbt
        self.loss, self.ignore_index = loss, float(ignore_index)


def _head_desc(spec, B, need_dx):
    L = len(spec.widths) - 1
    if L < 1 or L > _lib.HEAD_MAX_LAYERS or max(spec.widths) > _lib.HEAD_MAX_WIDTH:
        raise RuntimeError("openchem_b200: the fused MLP head supports 1..%d layers of width <= %d (got %s)" %
                           (_lib.HEAD_MAX_LAYERS, _lib.HEAD_MAX_WIDTH, spec.widths))
    if spec.training and L > 1 and B < 2:
        raise ValueError("Expected more than 1 value per channel when training, got input size [%d, %d]" % (B, spec.widths[1]))
    d = _lib.HeadDesc()
    d.B, d.n_layers = B, L
    for i, w in enumerate(spec.widths):
        d.width[i] = w
    for i, a in enumerate(spec.acts):
        d.act[i] = _lib.ACTS[a]
    d.training, d.loss, d.need_dx = int(spec.training), _lib.LOSSES[spec.loss], int(need_dx)
    d.eps, d.momentum, d.ignore_index = spec.eps, spec.momentum, spec.ignore_index
    return d


def _head_params(spec, Ws, bs, gammas, betas):
    p = _lib.HeadParams()
    for i in range(len(Ws)):
        p.W[i] = Ws[i].data_ptr()
        p.b[i] = bs[i].data_ptr() if bs[i] is not None else None
    for i in range(len(gammas)):
        p.gamma[i], p.beta[i] = gammas[i].data_ptr(), betas[i].data_ptr()
        p.running_mean[i], p.running_var[i] = spec.running_mean[i].data_ptr(), spec.running_var[i].data_ptr()
        p.num_batches_tracked[i] = spec.nbt[i].data_ptr() if spec.nbt[i] is not None else None
    return p


class HeadFn(torch.autograd.Function):
    """OpenChemMLP.forward (openchem_mlp.py:51-61), optionally fused with the criterion, as ONE autograd node.
    inputs: inp (B, widths[0]), target (B,T) or None, spec, then W_0..W_{L-1}, b_0..b_{L-1}, gamma_0.., beta_0.. (L-1 each).
    Returns predictions (spec.loss == 'none') or (loss scalar, predictions)."""

    @staticmethod
    def forward(ctx, inp, target, spec, *tensors):
        lib = _lib.load()
        _require_cuda(inp, "inp")
        L = len(spec.widths) - 1
        need_dx = bool(inp.requires_grad)
        inp = _f32(inp.detach()).contiguous()
        if inp.dim() != 2 or inp.shape[1] != spec.widths[0]:
            raise RuntimeError("openchem_b200: MLP input must be (B, %d); got %s" % (spec.widths[0], tuple(inp.shape)))
        ctx.sinks = _sinks_for(tensors)
        tensors = [None if t is None else t.detach().contiguous() for t in tensors]
        Ws, bs = tensors[0:L], tensors[L:2 * L]
        gammas, betas = tensors[2 * L:3 * L - 1], tensors[3 * L - 1:4 * L - 2]
        B, T = inp.shape[0], spec.widths[-1]
        fused = spec.loss != "none"
        if fused:
            _require_cuda(target, "target")
            target = _f32(target.detach()).contiguous()
            if tuple(target.shape) != (B, T):
                raise RuntimeError("openchem_b200: target must be (B, %d); got %s" % (T, tuple(target.shape)))
        named = [("layers[%d].weight" % i, Ws[i], torch.float32) for i in range(L)] + [("layers[%d].bias" % i, bs[i], torch.float32) for i in range(L)]
        for i in range(L - 1):
            named += [("bn[%d].weight" % i, gammas[i], torch.float32), ("bn[%d].bias" % i, betas[i], torch.float32),
                      ("bn[%d].running_mean" % i, spec.running_mean[i], torch.float32), ("bn[%d].running_var" % i, spec.running_var[i], torch.float32),
                      ("bn[%d].num_batches_tracked" % i, spec.nbt[i], torch.int64)]
        if fused:
            named.append(("target", target, torch.float32))
        _check_state(inp, named)
        with torch.cuda.device(inp.device):
            desc = _head_desc(spec, B, need_dx)
            params = _head_params(spec, Ws, bs, gammas, betas)
            sb, cb = ctypes.c_size_t(0), ctypes.c_size_t(0)
            _lib.check(lib.ocgcn_head_query_workspace(ctypes.byref(desc), ctypes.byref(sb), ctypes.byref(cb)))
            saved, scratch = _alloc(sb.value, inp.device), _alloc(cb.value, inp.device)
            pred = torch.empty((B, T), dtype=torch.float32, device=inp.device)
            loss = torch.empty((), dtype=torch.float32, device=inp.device) if fused else None
            _lib.check(lib.ocgcn_head_forward(ctypes.byref(desc), inp.data_ptr(), ctypes.byref(params),
                                              target.data_ptr() if fused else None, saved.data_ptr(), scratch.data_ptr(),
                                              pred.data_ptr(), loss.data_ptr() if fused else None, _stream_ptr(inp.device)))
            _count(lib)
        ctx.spec, ctx.cbytes, ctx.need_dx, ctx.fused = spec, cb.value, need_dx, fused
        ctx.none_mask = [t is None for t in tensors]
        ctx.save_for_backward(inp, pred, saved, *([target] if fused else []), *[t for t in tensors if t is not None])
        if fused:
            ctx.mark_non_differentiable(pred)
            return loss, pred
        return pred

    @staticmethod
    def backward(ctx, *grad_outputs):
        lib = _lib.load()
        spec = ctx.spec
        L = len(spec.widths) - 1
        inp, pred, saved, *rest = ctx.saved_tensors
        target = rest.pop(0) if ctx.fused else None
        it = iter(rest)
        tensors = [None if isnone else next(it) for isnone in ctx.none_mask]
        Ws, bs = tensors[0:L], tensors[L:2 * L]
        gammas, betas = tensors[2 * L:3 * L - 1], tensors[3 * L - 1:4 * L - 2]
        grad = _f32(grad_outputs[0]).contiguous()
        with torch.cuda.device(inp.device):
            desc = _head_desc(spec, inp.shape[0], ctx.need_dx)
            params = _head_params(spec, Ws, bs, gammas, betas)
            grads = _lib.HeadGrads()
            sk = ctx.sinks
            gW = [_grad_buffer(w, sk, i) for i, w in enumerate(Ws)]
            gb = [None if b is None else _grad_buffer(b, sk, L + i) for i, b in enumerate(bs)]
            gg = [_grad_buffer(t, sk, 2 * L + i) for i, t in enumerate(gammas)]
            gbt = [_grad_buffer(t, sk, 3 * L - 1 + i) for i, t in enumerate(betas)]
            for i in range(L):
                grads.dW[i] = gW[i].data_ptr()
                grads.db[i] = gb[i].data_ptr() if gb[i] is not None else None
            for i in range(L - 1):
                grads.dgamma[i], grads.dbeta[i] = gg[i].data_ptr(), gbt[i].data_ptr()
            gx = torch.empty_like(inp) if ctx.need_dx else None
            scratch = _alloc(ctx.cbytes, inp.device)
            _lib.check(lib.ocgcn_head_backward(ctypes.byref(desc), grad.data_ptr(), inp.data_ptr(), pred.data_ptr(),
                                               target.data_ptr() if target is not None else None, ctypes.byref(params),
                                               ctypes.byref(grads), gx.data_ptr() if gx is not None else None, saved.data_ptr(),
                                               scratch.data_ptr(), _stream_ptr(inp.device)))
            _count(lib)
        return (gx, None, None, *gW, *gb, *gg, *gbt)


# ---------------------------------------------------------------------------------------------------------------------
# tanh + neighbour max-pool as a stand-alone op — ocgcn_pool_forward / ocgcn_pool_backward
# ---------------------------------------------------------------------------------------------------------------------
class PoolFn(torch.autograd.Function):
    """pooled[b,i,c] = max_j adj[b,i,j] * tanh(y[b,j,c]) (first index on ties): the tanh + x.repeat/view/expand/max(dim=2) of
    edge_attention_encoder.py:54-61 / gcn_encoder.py:43-50 without the (B,N,N,D) temporaries. Differentiable w.r.t. y."""

    @staticmethod
    def forward(ctx, y, adj):
        lib = _lib.load()
        _require_cuda(y, "y")
        _require_cuda(adj, "adj")
        if adj.requires_grad:
            raise RuntimeError("openchem_b200: gradients w.r.t. the adjacency are not supported (no reference caller needs them)")
        y = _f32(y.detach()).contiguous()
        adj = _f32(adj)
        B, N, D = y.shape
        if tuple(adj.shape) != (B, N, N) or adj.device != y.device:
            raise RuntimeError("openchem_b200: adj must be (B,N,N) on the device of y (B,N,D); got %s and %s" % (tuple(adj.shape), tuple(y.shape)))
        pooled = torch.empty_like(y)
        arg = torch.empty((B, N, D), dtype=torch.uint8, device=y.device)
        strides = (ctypes.c_int64 * 3)(*adj.stride())
        with torch.cuda.device(y.device):
            _lib.check(lib.ocgcn_pool_forward(B, N, D, y.data_ptr(), adj.data_ptr(), strides, pooled.data_ptr(), arg.data_ptr(),
                                              _stream_ptr(y.device)))
            _count(lib)
        ctx.save_for_backward(y, adj, arg)
        return pooled

    @staticmethod
    def backward(ctx, grad_pooled):
        lib = _lib.load()
        y, adj, arg = ctx.saved_tensors
        B, N, D = y.shape
        grad_pooled = _f32(grad_pooled).contiguous()
        grad_y = torch.empty_like(y)
        strides = (ctypes.c_int64 * 3)(*adj.stride())
        with torch.cuda.device(y.device):
            _lib.check(lib.ocgcn_pool_backward(B, N, D, y.data_ptr(), adj.data_ptr(), strides, arg.data_ptr(), grad_pooled.data_ptr(),
                                               grad_y.data_ptr(), _stream_ptr(y.device)))
            _count(lib)
        return grad_y, None


def tanh_neighbour_max_pool(y, adj):
    """Public name of PoolFn: tanh, then the neighbour max-pool over `adj`."""
    return PoolFn.apply(y, adj)
