"""Step-level runtime for the data-parallel GraphCNN train step.

One train step = zero grads -> model forward -> criterion -> backward -> gradient all-reduce (NCCL, average) ->
optimizer step. The reference drives these from Python every iteration (openchem/models/openchem_model.py:102-121, under
torch DDP, run.py:235-236); at the 1-3 ms steps of the B200 path that host work and DDP's per-iteration bookkeeping
show up on the critical path. `TrainStep` captures the whole step — including the all-reduce — in ONE CUDA graph:

* gradients live in one flat fp32 buffer (every `p.grad` is a view), so the exchange is a single all-reduce of 0.4 MB issued
  inside the graph right after backward (NCCL; symmetric-memory one-shot / NVLS multimem variants over NVLink peer memory are
  selectable, see `_symmetric_flat_grad`);
* BatchNorm statistics stay per rank (plain BatchNorm1d semantics, as under the reference's DDP); running buffers are not
  re-broadcast every step — rank 0's are the ones a checkpoint written by rank 0 holds, as in the reference;
* inputs are read from static device buffers (`.inputs`); copy a new batch into them (any stream-ordered copy) and replay.

Nothing here is needed to use the layer/encoder modules; the reference's own `fit` loop keeps working unchanged.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class DelayedLoss:
    """Reads every step's loss on the host without stalling the device: the device -> host copy of step t's scalar is
    enqueued behind step t, and its VALUE is collected while step t+1 is already running (the reference's `fit` loop calls
    `.item()` right after `backward`, openchem/models/openchem_model.py:159-163, which idles the GPU for the host's
    launch latency every iteration). `push` returns the previous step's loss (None on the first call); `flush` the last."""

    def __init__(self, device, depth=2):
        self.host = torch.zeros(depth, dtype=torch.float32).pin_memory()
        self.events = [torch.cuda.Event() for _ in range(depth)]
        self.device, self.depth, self.n = torch.device(device), depth, 0

    def _collect(self, k):
        self.events[k].synchronize()
        return float(self.host[k])

    def push(self, loss):
        prev = self._collect((self.n - 1) % self.depth) if self.n > 0 else None
        k = self.n % self.depth
        self.host[k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        self.events[k].record(torch.cuda.current_stream(self.device))
        self.n += 1
        return prev

    def flush(self):
        return self._collect((self.n - 1) % self.depth) if self.n > 0 else None


class TrainStep:
    def __init__(self, model, criterion, make_optimizer, example_batch, use_graph=True, warmup=3, loss_fn=None, exchange=True,
                 allreduce=None):
        """model((x, adj)) -> predictions; criterion(pred, y) -> scalar; make_optimizer(params) -> torch optimizer that is
        CUDA-graph capturable (e.g. Adam(fused=True, capturable=True)); example_batch = (x, adj, y) device tensors.
        loss_fn(model, (x, adj), y) -> scalar, optional: replaces criterion(model(inp), y), e.g. to evaluate the criterion
        inside the MLP head's last kernel (openchem_b200.modules.mlp.openchem_mlp.fused_head_loss).
        Construction runs `warmup` real steps and the capture on `example_batch`, then restores the model / optimizer state it was
        given (see the snapshot note below): the first __call__ is training step 1."""
        self.model, self.criterion, self.loss_fn = model, criterion, loss_fn
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if not exchange:      # a purely local step (no gradient all-reduce, no initial broadcast): single-GPU timing inside a multi-rank job
            self.world = 1
        dev = example_batch[0].device
        self.params = [p for p in model.parameters() if p.requires_grad]
        numel = sum(p.numel() for p in self.params)
        self.exchange_kind, self._group_name = "none", None
        self.flat_grad = None
        if self.world > 1:
            self.flat_grad = self._symmetric_flat_grad(numel, dev, allreduce)
        if self.flat_grad is None:
            self.flat_grad = torch.zeros(numel, dtype=torch.float32, device=dev)
            if self.world > 1:
                self.exchange_kind = "nccl"
        off = 0
        self.views = []
        from . import functional
        for p in self.params:
            n = p.numel()
            p.grad = self.flat_grad[off:off + n].view_as(p)
            self.views.append(p.grad)
            functional.grad_sinks[p] = p.grad     # the package's backward kernels write the gradient straight into this view
            off += n
        self.opt = make_optimizer(self.params)
        self.inputs = tuple(torch.empty_like(t) for t in example_batch)
        for dst, src in zip(self.inputs, example_batch):
            dst.copy_(src)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.graph, self.graph_error = None, None
        if self.world > 1:   # identical initial weights on every rank (DDP does this at construction)
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=0)
        # Warm-up and capture must not advance training: the warm-up bodies below are real optimizer steps on `example_batch`. The
        # model's parameters and buffers (BatchNorm running statistics, num_batches_tracked) are snapshotted here and restored IN
        # PLACE afterwards, and the optimizer's state tensors are zeroed in place (the captured graph holds their addresses), so a
        # freshly built TrainStep starts from exactly the state it was given - a reloaded checkpoint gets no hidden extra updates.
        snapshot = {k: v.detach().clone() for k, v in model.state_dict().items()}
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(int(warmup), 1)):
                self._body()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        if use_graph:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._body()
                self.graph = g
            except Exception as exc:  # noqa: BLE001 - capture is an optimisation; the eager body is the same work
                self.graph, self.graph_error = None, repr(exc)[:300]
                torch.cuda.synchronize(dev)
        with torch.no_grad():
            for k, v in model.state_dict().items():
                v.copy_(snapshot[k])
            for st in self.opt.state.values():
                for v in st.values():
                    if isinstance(v, torch.Tensor):
                        v.zero_()
            self.flat_grad.zero_()
            self.loss.zero_()
        torch.cuda.synchronize(dev)

    def _body(self):
        from . import functional
        x, adj, y = self.inputs
        functional.sink_scope["on"] = True        # (decided at forward time; see functional.grad_sinks)
        try:
            loss = self.loss_fn(self.model, (x, adj), y) if self.loss_fn is not None else self.criterion(self.model((x, adj)), y)
        finally:
            functional.sink_scope["on"] = False
        # torch.autograd.grad instead of .backward(): no AccumulateGrad node runs. Gradients the package's kernels produced are already
        # in the flat buffer (the returned tensor IS the view); anything else (torch modules in the model) is copied, unused -> zero
        grads = torch.autograd.grad(loss, self.params, allow_unused=True)
        for view, g in zip(self.views, grads):
            if g is None:
                view.zero_()
            elif g.data_ptr() != view.data_ptr():
                view.copy_(g)
        if self.world > 1:
            self._exchange()
        self.opt.step()
        self.loss.copy_(loss.detach())

    def _symmetric_flat_grad(self, numel, dev, allreduce):
        """The gradient exchange is ONE latency-bound 0.4 MB all-reduce per step. Over NVLink / NVSwitch it does not need NCCL's ring:
        with the flat gradient in symmetric (peer-mapped) memory every rank can read all peers' buffers directly (`one_shot`: sum in
        rank order, identical on every rank) or let the switch reduce (`multimem`, NVLS). `allreduce` / $OCB200_ALLREDUCE selects
        'nccl' (default) | 'one_shot' | 'multimem'. Measured at 2 B200 (c4, 30 steps): 1.223 ms per step with NCCL, 1.233 one_shot,
        1.228 multimem, against 1.192 without any exchange - the ~30 us are the per-step wait for the slower rank (every exchange
        is a synchronisation point), not transfer time, so the direct-access variants buy nothing here; they stay selectable and
        covered by tests/test_gpu_dist.py."""
        import os
        kind = (allreduce or os.environ.get("OCB200_ALLREDUCE", "nccl")).lower()
        if kind in ("nccl", "auto"):
            return None
        try:
            import torch.distributed._symmetric_memory as symm
            group = dist.group.WORLD
            buf = symm.empty(numel, dtype=torch.float32, device=dev)
            symm.rendezvous(buf, group.group_name)
            buf.zero_()
            self._group_name = group.group_name
            self.exchange_kind = "multimem" if kind == "multimem" else "one_shot"
            self._reduced = torch.empty(numel, dtype=torch.float32, device=dev)
            return buf
        except Exception as exc:  # noqa: BLE001 - any failure of the optional fast path falls back to the NCCL all-reduce
            raise RuntimeError("openchem_b200.TrainStep: allreduce=%r needs torch symmetric memory over NVLink peers: %r" % (kind, exc))

    def _exchange(self):
        if self.exchange_kind == "nccl":
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.AVG)
        elif self.exchange_kind == "multimem":
            torch.ops.symm_mem.multimem_all_reduce_(self.flat_grad, "sum", self._group_name)
            self.flat_grad.mul_(1.0 / self.world)
        else:   # one_shot: every rank sums all peers' buffers (barriers inside the op keep the buffers stable while they are read)
            torch.ops.symm_mem.one_shot_all_reduce_out(self.flat_grad, "sum", self._group_name, self._reduced)
            torch.mul(self._reduced, 1.0 / self.world, out=self.flat_grad)

    def load(self, x, adj, y):
        """Stream-ordered copy of a batch into the static input buffers."""
        for dst, src in zip(self.inputs, (x, adj, y)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)

    def load_compact(self, batch):
        """Feed a device CompactGraphBatch (openchem_b200.compact); only the bit words crossed the host link.
        A step built on a compact example batch (`example_batch = (x_bits, adj_bits, labels)`, int32 words) keeps the words as
        its static inputs and the captured kernels read them directly (ocgcn_encoder_forward_compact): three small copies.
        A step built on dense tensors gets the batch expanded straight into its static dense inputs (no staging copy)."""
        if self.inputs[0].dtype == torch.int32:
            self.inputs[0].copy_(batch.x_bits, non_blocking=True)
            self.inputs[1].copy_(batch.adj_bits, non_blocking=True)
        else:
            batch.expand(out=(self.inputs[0], self.inputs[1]))
        if batch.labels is not None:
            self.inputs[2].copy_(batch.labels, non_blocking=True)

    def __call__(self, x=None, adj=None, y=None):
        if x is not None:
            self.load(x, adj, y)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        return self.loss
