"""Host side of the compact batch format (openchem_b200/compact.py): bit layout, round trips, rejection of non-binary
matrices, the DataLoader collate function over the reference's sample dicts. No GPU needed."""
import numpy as np
import pytest
import torch

from openchem_b200 import compact
from openchem_b200.synthetic import make_graph_batch, make_labels


def _bits_by_hand(dense):
    """Literal restatement of the layout in include/ocgcn.h: bit (c & 31) of word c >> 5."""
    *lead, C = dense.shape
    W = (C + 31) // 32
    out = np.zeros(lead + [W], dtype=np.uint32)
    for idx in np.ndindex(*lead):
        for c in range(C):
            if dense[idx + (c,)] == 1:
                out[idx + (c >> 5,)] |= np.uint32(1) << np.uint32(c & 31)
    return out


@pytest.mark.parametrize("C", [1, 7, 31, 32, 33, 64, 75, 132])
def test_pack_matches_layout_and_round_trips(C):
    rng = np.random.RandomState(C)
    dense = (rng.rand(3, 5, C) < 0.3).astype(np.float32)
    bits = compact.pack_bits_host(dense)
    assert bits.dtype == np.int32 and bits.shape == (3, 5, (C + 31) // 32)
    assert np.array_equal(bits.view(np.uint32), _bits_by_hand(dense))
    assert np.array_equal(compact.unpack_bits_host(bits, C), dense)


def test_non_binary_is_rejected():
    x = np.zeros((1, 4, 8), np.float32)
    adj = np.eye(4, dtype=np.float32)[None]
    compact.pack_graph_batch(x, adj)
    bad = adj.copy()
    bad[0, 1, 2] = 0.5
    with pytest.raises(ValueError):
        compact.pack_graph_batch(x, bad)
    badx = x.copy()
    badx[0, 0, 0] = 2.0
    with pytest.raises(ValueError):
        compact.pack_graph_batch(badx, adj)
    with pytest.raises(ValueError):
        compact.pack_graph_batch(x, np.eye(5, dtype=np.float32)[None])     # adj does not match x


def test_synthetic_batch_round_trip_and_size():
    g = make_graph_batch(16, 50, 33, seed=3, occupancy="molecular")
    cb = compact.pack_graph_batch(g["x"], g["adj"], make_labels(16))
    assert cb.batch_size == 16 and cb.n_atoms == 50 and cb.n_features == 33
    assert np.array_equal(compact.unpack_bits_host(cb.x_bits.numpy(), 33), g["x"])
    assert np.array_equal(compact.unpack_bits_host(cb.adj_bits.numpy(), 50), g["adj"])
    dense_bytes = g["x"].nbytes + g["adj"].nbytes
    assert cb.nbytes() - cb.labels.numel() * 4 == 16 * 50 * (2 + 2) * 4 and dense_bytes / cb.nbytes() > 15


def test_collate_over_reference_sample_dicts():
    g = make_graph_batch(6, 20, 33, seed=1, occupancy="molecular")
    g["y"] = make_labels(6)
    samples = [{"adj_matrix": g["adj"][i], "node_feature_matrix": g["x"][i], "labels": g["y"][i]} for i in range(6)]
    cb = compact.compact_collate(samples)
    assert torch.equal(cb.labels, torch.from_numpy(g["y"]))
    assert np.array_equal(compact.unpack_bits_host(cb.adj_bits.numpy(), 20), g["adj"])
    # usable as a torch DataLoader collate_fn
    dl = torch.utils.data.DataLoader(samples, batch_size=3, collate_fn=compact.compact_collate)
    parts = list(dl)
    assert len(parts) == 2 and parts[1].batch_size == 3
    assert np.array_equal(compact.unpack_bits_host(parts[1].x_bits.numpy(), 33), g["x"][3:])


def test_expand_refuses_host_tensors():
    g = make_graph_batch(2, 8, 16, seed=0)
    cb = compact.pack_graph_batch(g["x"], g["adj"])
    with pytest.raises(RuntimeError):
        cb.expand()
