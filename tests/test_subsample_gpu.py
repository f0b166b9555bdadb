"""
GPU: minibatch index draws (tq_subsample; the reference leaves them to pyro.plate [third party]: randperm(size)[:n],
models/cosmos.py:194-208).  No RNG-stream parity is possible with Pyro, so the properties are checked: n distinct
indices in range, a function of (seed, stream, step) only, every index equally likely, every position of the ordered
sample equally likely to hold a given index -- for the ranked form (axes up to 16384) and the serial fallback beyond.
"""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def draw(n_total, n_pick, seed, step, stream, perm=None):
    from tapqir_b200 import _lib

    lib, p = _lib.load(), _lib.ptr
    dev = torch.device("cuda")
    state = torch.tensor([step], dtype=torch.int64, device=dev)
    perm = torch.arange(n_total, dtype=torch.int32, device=dev) if perm is None else perm
    out = torch.full((n_pick,), -1, dtype=torch.int32, device=dev)
    _lib.check(lib.tq_subsample(n_total, n_pick, seed, p(state), stream, p(perm), p(out), _lib.stream_ptr(dev)), "tq_subsample")
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("n_total,n_pick", [(1000, 512), (100, 10), (5, 5), (7, 1), (16384, 300), (20000, 64)])
def test_indices_are_distinct_in_range_and_reproducible(n_total, n_pick):
    a = draw(n_total, n_pick, 3, 11, 1)
    assert a.min() >= 0 and a.max() < n_total and len(set(a.tolist())) == n_pick
    assert (draw(n_total, n_pick, 3, 11, 1) == a).all()                 # same (seed, stream, step): same draw
    if n_pick < n_total or n_total > 2:
        others = [draw(n_total, n_pick, 3, 12, 1), draw(n_total, n_pick, 3, 11, 2), draw(n_total, n_pick, 4, 11, 1)]
        assert sum((o != a).any() for o in others) >= 2                 # step, stream and seed all enter


def test_every_index_and_every_position_is_equally_likely():
    n_total, n_pick, reps = 40, 10, 4000
    hits = np.zeros(n_total)
    first = np.zeros(n_total)
    for step in range(reps):
        a = draw(n_total, n_pick, 0, step, 1)
        hits[a] += 1
        first[a[0]] += 1
    p = n_pick / n_total
    assert np.abs(hits / reps - p).max() < 5 * np.sqrt(p * (1 - p) / reps)
    assert np.abs(first / reps - 1 / n_total).max() < 5 * np.sqrt((1 / n_total) / reps)


@pytest.mark.parametrize("sizes", [((100, 10), (1000, 512)), ((7, 3), (5, 5)), ((16384, 40), (300, 300))])
def test_pair_launch_draws_the_same_indices_as_two_launches(sizes):
    """tq_subsample_pair (both axes of a step in one launch) = two tq_subsample calls, index for index."""
    from tapqir_b200 import _lib

    lib, p = _lib.load(), _lib.ptr
    dev = torch.device("cuda")
    (n0, k0), (n1, k1) = sizes
    assert lib.tq_subsample_pair_supported(n0, n1) == 1 and lib.tq_subsample_pair_supported(20000, 10) == 0
    state = torch.tensor([17], dtype=torch.int64, device=dev)
    out0 = torch.full((k0,), -1, dtype=torch.int32, device=dev)
    out1 = torch.full((k1,), -1, dtype=torch.int32, device=dev)
    _lib.check(lib.tq_subsample_pair(n0, k0, 2, p(out0), n1, k1, 1, p(out1), 5, p(state), _lib.stream_ptr(dev)), "tq_subsample_pair")
    torch.cuda.synchronize()
    assert (out0.cpu().numpy() == draw(n0, k0, 5, 17, 2)).all() and (out1.cpu().numpy() == draw(n1, k1, 5, 17, 1)).all()
