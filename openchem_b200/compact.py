"""Compact batch format for the GraphCNN path (C ABI: ocgcn_expand_compact / ocgcn_pack_compact).

The reference's `Graph2Label.cast_inputs` (openchem/models/Graph2Label.py:41-57) uploads dense fp32 `adj_matrix (B,N,N)` and
`node_feature_matrix (B,N,F0)` with a synchronous `.to('cuda')` every iteration: 4N^2 + 4N*F0 bytes per molecule (32 KiB at
N = F0 = 64). For the 0/1 matrices its data layer produces (one-hot attribute groups, openchem/utils/graph.py:57-63; bonds
plus self loops, :80-84) the same batch is N*ceil(N/32) + N*ceil(F0/32) 32-bit words (1 KiB): bit (c & 31) of word
`bits[mol, row, c >> 5]` is set iff `dense[mol, row, c] == 1`.

* `pack_graph_batch` / `compact_collate` build the words on the HOST (numpy) from the reference's sample dicts;
* `CompactGraphBatch.expand()` rebuilds the dense fp32 tensors ON THE DEVICE (bit-exact), optionally gathering molecules
  by index, so a whole data set can stay resident in HBM in compact form (`ResidentCompactDataset`);
* `GraphCNNEncoder.forward` accepts a `CompactGraphBatch` in place of the `(x, adj)` tuple.

Non-binary matrices (e.g. the soft adjacencies of GraphEdgeAttentionEncoder) are rejected: they must use the dense format.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib


def words(c):
    return (int(c) + 31) // 32


def pack_bits_host(dense):
    """(..., C) array of exact 0/1 values -> (..., ceil(C/32)) int32 words (bit c&31 of word c>>5). Raises on other values."""
    a = np.asarray(dense)
    nz = a != 0
    if not np.array_equal(nz, a == 1):
        raise ValueError("openchem_b200.compact: matrix has entries other than 0 and 1; use the dense (x, adj) format")
    C = a.shape[-1]
    W = words(C)
    by = np.packbits(nz, axis=-1, bitorder="little")
    pad = W * 4 - by.shape[-1]
    if pad:
        by = np.concatenate([by, np.zeros(by.shape[:-1] + (pad,), dtype=np.uint8)], axis=-1)
    return np.ascontiguousarray(by).view("<u4").view(np.int32).reshape(a.shape[:-1] + (W,))


def unpack_bits_host(bits, C):
    """Inverse of pack_bits_host on the host (tests and tools; the product path expands on the device)."""
    w = np.ascontiguousarray(np.asarray(bits)).view(np.uint32).astype("<u4")
    by = w.view(np.uint8).reshape(w.shape[:-1] + (w.shape[-1] * 4,))
    return np.unpackbits(by, axis=-1, bitorder="little")[..., :C].astype(np.float32)


class CompactGraphBatch:
    """x_bits (B,N,ceil(F0/32)) int32, adj_bits (B,N,ceil(N/32)) int32, optional labels; host or device tensors."""

    def __init__(self, x_bits, adj_bits, n_features, labels=None):
        if x_bits.dtype != torch.int32 or adj_bits.dtype != torch.int32:
            raise TypeError("openchem_b200.compact: bit words are int32 tensors")
        B, N, XW = x_bits.shape
        if tuple(adj_bits.shape) != (B, N, words(N)) or XW != words(n_features):
            raise ValueError("openchem_b200.compact: inconsistent shapes x_bits %s adj_bits %s for F0=%d" %
                             (tuple(x_bits.shape), tuple(adj_bits.shape), n_features))
        self.x_bits, self.adj_bits, self.labels = x_bits, adj_bits, labels
        self.n_features = int(n_features)

    # -- shape / placement -------------------------------------------------------------------------------------------
    @property
    def batch_size(self):
        return self.x_bits.shape[0]

    @property
    def n_atoms(self):
        return self.x_bits.shape[1]

    @property
    def device(self):
        return self.x_bits.device

    def nbytes(self):
        n = self.x_bits.numel() * 4 + self.adj_bits.numel() * 4
        return n + (self.labels.numel() * self.labels.element_size() if self.labels is not None else 0)

    def tensors(self):
        return tuple(t for t in (self.x_bits, self.adj_bits, self.labels) if t is not None)

    def _map(self, fn):
        return CompactGraphBatch(fn(self.x_bits), fn(self.adj_bits), self.n_features,
                                 None if self.labels is None else fn(self.labels))

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())

    def to(self, device, non_blocking=False):
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    # -- device expansion --------------------------------------------------------------------------------------------
    def expand(self, out=None, index=None):
        """-> dense fp32 (x (B,N,F0), adj (B,N,N)) on the device, written by ocgcn_expand_compact on the current stream.
        out: optional preallocated contiguous (x, adj) pair (e.g. the static inputs of a captured step).
        index: optional int32 device tensor (B',) selecting molecules -> outputs have batch B'."""
        if not self.x_bits.is_cuda:
            raise RuntimeError("openchem_b200.compact: expand() runs on the device (there is no CPU path); call .to('cuda') first")
        xb, ab = self.x_bits.contiguous(), self.adj_bits.contiguous()
        N, F0 = self.n_atoms, self.n_features
        oob = None
        if index is not None:
            index = self._checked_index(index)
            oob = torch.empty(1, dtype=torch.int32, device=self.device)
        B = self.batch_size if index is None else index.shape[0]
        if out is None:
            x = torch.empty((B, N, F0), dtype=torch.float32, device=self.device)
            adj = torch.empty((B, N, N), dtype=torch.float32, device=self.device)
        else:
            x, adj = out
            if tuple(x.shape) != (B, N, F0) or tuple(adj.shape) != (B, N, N) or not x.is_contiguous() or not adj.is_contiguous() \
                    or x.dtype != torch.float32 or adj.dtype != torch.float32:
                raise ValueError("openchem_b200.compact: out must be contiguous fp32 (B,N,F0) and (B,N,N)")
        lib = _lib.load()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(lib.ocgcn_expand_compact(B, N, F0, xb.data_ptr(), ab.data_ptr(), index.data_ptr() if index is not None else None,
                                                self.batch_size, oob.data_ptr() if oob is not None else None,
                                                x.data_ptr(), adj.data_ptr(), stream))
        from . import functional
        functional.launch_counter["n"] += lib.ocgcn_last_launch_count()
        self.last_index_oob = oob      # device flag of this gather (1 = some index was out of range and expanded to zeros)
        return x, adj

    def _checked_index(self, index):
        """-> 1-D int32 CUDA index. Host indices (list / numpy / CPU tensor: the usual case, a sampler's permutation) are
        bounds-checked on the host before the upload; a CUDA index is checked by the kernel itself, which never dereferences
        an out-of-range entry and raises `last_index_oob` (read it with .item() when a synchronisation is acceptable)."""
        n = self.batch_size
        if not (isinstance(index, torch.Tensor) and index.is_cuda):
            host = torch.as_tensor(index)
            if host.dim() != 1 or host.dtype not in (torch.int32, torch.int64, torch.int16, torch.uint8, torch.int8):
                raise TypeError("openchem_b200.compact: index must be a 1-D integer tensor")
            if host.numel() and (int(host.min()) < 0 or int(host.max()) >= n):
                raise IndexError("openchem_b200.compact: index out of range for a data set of %d molecules (min %d, max %d)" %
                                 (n, int(host.min()), int(host.max())))
            return host.to(torch.int32).to(self.device).contiguous()
        if index.dim() != 1 or index.dtype not in (torch.int32, torch.int64):
            raise TypeError("openchem_b200.compact: index must be a 1-D int32 / int64 CUDA tensor")
        if index.dtype == torch.int64:      # values that do not fit in int32 must not wrap into range
            index = index.clamp(min=-1, max=2 ** 31 - 1).to(torch.int32)
        return index.to(self.device).contiguous()


def pack_graph_batch(x, adj, labels=None):
    """Host packing of one reference-format batch: x (B,N,F0), adj (B,N,N) numpy arrays / CPU tensors of exact 0/1 values."""
    xn = x.numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    an = adj.numpy() if isinstance(adj, torch.Tensor) else np.asarray(adj)
    if xn.ndim != 3 or an.ndim != 3 or an.shape[0] != xn.shape[0] or an.shape[1] != xn.shape[1] or an.shape[2] != xn.shape[1]:
        raise ValueError("openchem_b200.compact: expected x (B,N,F0) and adj (B,N,N); got %s and %s" % (xn.shape, an.shape))
    lab = None
    if labels is not None:
        lab = labels if isinstance(labels, torch.Tensor) else torch.from_numpy(np.asarray(labels))
    return CompactGraphBatch(torch.from_numpy(pack_bits_host(xn)), torch.from_numpy(pack_bits_host(an)), xn.shape[2], lab)


def pack_graph_batch_device(x, adj, labels=None):
    """Device packing (ocgcn_pack_compact) of dense CUDA tensors with arbitrary strides; raises on non-binary entries."""
    if not x.is_cuda or not adj.is_cuda:
        raise RuntimeError("openchem_b200.compact: pack_graph_batch_device needs CUDA tensors")
    x = x if x.dtype == torch.float32 else x.to(torch.float32)
    adj = adj if adj.dtype == torch.float32 else adj.to(torch.float32)
    B, N, F0 = x.shape
    xb = torch.empty((B, N, words(F0)), dtype=torch.int32, device=x.device)
    ab = torch.empty((B, N, words(N)), dtype=torch.int32, device=x.device)
    flag = torch.empty((1,), dtype=torch.int32, device=x.device)
    xs = (ctypes.c_int64 * 3)(*x.stride())
    as_ = (ctypes.c_int64 * 3)(*adj.stride())
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.ocgcn_pack_compact(B, N, F0, x.data_ptr(), xs, adj.data_ptr(), as_, xb.data_ptr(), ab.data_ptr(), flag.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
    if int(flag.item()):
        raise ValueError("openchem_b200.compact: matrix has entries other than 0 and 1; use the dense (x, adj) format")
    return CompactGraphBatch(xb, ab, F0, labels)


def compact_collate(samples):
    """DataLoader `collate_fn` over the reference's sample dicts ({'adj_matrix','node_feature_matrix','labels'},
    openchem/data/graph_data_layer.py:104-116) -> host CompactGraphBatch (labels as float32)."""
    x = np.stack([np.asarray(s["node_feature_matrix"]) for s in samples])
    adj = np.stack([np.asarray(s["adj_matrix"]) for s in samples])
    labels = None
    if "labels" in samples[0]:
        labels = torch.from_numpy(np.stack([np.asarray(s["labels"], dtype=np.float32) for s in samples]))
    return pack_graph_batch(x, adj, labels)


class ResidentCompactDataset:
    """A whole data set kept in device memory in compact form (1 KiB per molecule at N = F0 = 64: 10^8 molecules fit in the
    180 GB of one B200). `batch(index)` expands the selected molecules straight into dense fp32 batch tensors: the per-step
    host -> device traffic is the index vector (4 bytes per molecule)."""

    def __init__(self, compact, device):
        self.data = compact.to(device)

    def __len__(self):
        return self.data.batch_size

    def batch(self, index, out=None):
        index = self.data._checked_index(index)
        x, adj = self.data.expand(out=out, index=index)
        labels = None
        if self.data.labels is not None:      # (clamped: an out-of-range entry was already flagged / expanded to zeros above)
            labels = self.data.labels.index_select(0, index.long().clamp(0, len(self) - 1))
        return x, adj, labels


class CompactGraphDataset(torch.utils.data.Dataset):
    """A featurised graph data set cached on the HOST in compact form ("cached / pre-featurised
    GraphDataset" - the reference's GraphDataset re-parses and re-featurises every molecule in every __getitem__,
    openchem/data/graph_data_layer.py:104-116).

    Built ONCE from any map-style data set that yields the reference's sample dicts (or from dense arrays); afterwards a batch
    is a row gather of 32-bit words: `collate(indices)` writes the selected molecules into one of a few pinned staging buffers
    and returns a pinned CompactGraphBatch ready for a non-blocking upload (1 KiB per molecule at N = F0 = 64 instead of 32 KiB).
    `__getitem__` keeps the reference's dict layout (dense fp32), so the object also works with run.py's stock DataLoader."""

    def __init__(self, x_bits, adj_bits, n_features, labels=None, pinned_slots=4):
        self.x_bits, self.adj_bits = torch.as_tensor(x_bits), torch.as_tensor(adj_bits)
        self.labels = None if labels is None else torch.as_tensor(labels, dtype=torch.float32)
        self.n_features = int(n_features)
        CompactGraphBatch(self.x_bits, self.adj_bits, self.n_features, self.labels)      # shape / dtype validation
        self.max_size = int(self.x_bits.shape[1])
        self._slots, self._k, self._n_slots = {}, 0, int(pinned_slots)

    @classmethod
    def from_dense(cls, x, adj, labels=None, **kw):
        cb = pack_graph_batch(x, adj, labels)
        return cls(cb.x_bits, cb.adj_bits, cb.n_features, cb.labels, **kw)

    @classmethod
    def from_dataset(cls, dataset, chunk=1024, **kw):
        """Featurise every sample of `dataset` once (e.g. the reference's GraphDataset) and keep only the packed words."""
        parts = []
        for lo in range(0, len(dataset), chunk):
            parts.append(compact_collate([dataset[i] for i in range(lo, min(len(dataset), lo + chunk))]))
        lab = None if parts[0].labels is None else torch.cat([p.labels for p in parts])
        return cls(torch.cat([p.x_bits for p in parts]), torch.cat([p.adj_bits for p in parts]), parts[0].n_features, lab, **kw)

    def __len__(self):
        return self.x_bits.shape[0]

    def __getitem__(self, i):
        s = {"adj_matrix": unpack_bits_host(self.adj_bits[i].numpy(), self.max_size),
             "node_feature_matrix": unpack_bits_host(self.x_bits[i].numpy(), self.n_features)}
        if self.labels is not None:
            s["labels"] = self.labels[i].numpy()
        return s

    def collate(self, indices):
        """Gather the molecules `indices` (any 1-D integer sequence) into a pinned host CompactGraphBatch. The staging buffers
        rotate through `pinned_slots` sets: a returned batch stays valid for the next pinned_slots - 1 calls."""
        idx = torch.as_tensor(indices, dtype=torch.int64)
        if idx.dim() != 1 or (idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= len(self))):
            raise IndexError("openchem_b200.compact: batch indices out of range for a data set of %d molecules" % len(self))
        n = idx.numel()
        key = (n, self._k % self._n_slots)
        self._k += 1
        if key not in self._slots:
            pin = torch.cuda.is_available()
            mk = lambda src: torch.empty((n,) + tuple(src.shape[1:]), dtype=src.dtype, pin_memory=pin)  # noqa: E731
            self._slots[key] = (mk(self.x_bits), mk(self.adj_bits), None if self.labels is None else mk(self.labels))
        xs, as_, ls = self._slots[key]
        torch.index_select(self.x_bits, 0, idx, out=xs)
        torch.index_select(self.adj_bits, 0, idx, out=as_)
        if ls is not None:
            torch.index_select(self.labels, 0, idx, out=ls)
        return CompactGraphBatch(xs, as_, self.n_features, ls)
