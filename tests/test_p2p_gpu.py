"""
GPU: the peer-memory sum kernels (csrc/p2p_allreduce.cu) driven through the C ABI on ONE device -- two "ranks" living
in one process, each with its own buffer and the pointer table both kernels take -- over several calls (both parities of
the double buffering).  The multi-process path (CUDA IPC handles through the process group) is exercised by
profiles/multigpu_check.py and bench.py --gpus N.
"""

import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_push_wait_sum_two_ranks_in_one_process():
    from tapqir_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda")
    bufs = []
    for _ in range(2):
        b, h = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        _lib.check(lib.tq_p2p_alloc(ctypes.byref(b), ctypes.cast(h, ctypes.c_void_p)), "tq_p2p_alloc")
        bufs.append(b.value)
    peers = torch.tensor(bufs, dtype=torch.int64, device=dev)
    st = _lib.stream_ptr(dev)
    n = 18
    g = torch.Generator().manual_seed(0)
    try:
        for call in range(5):
            vals = [torch.randn(n, generator=g, dtype=torch.float64).to(dev) for _ in range(2)]
            outs = [torch.zeros(n, dtype=torch.float64, device=dev) for _ in range(2)]
            for r in range(2):
                _lib.check(lib.tq_p2p_push(_lib.ptr(vals[r]), n, r, 2, _lib.ptr(peers), st), "tq_p2p_push")
            for r in range(2):
                _lib.check(lib.tq_p2p_wait_sum(ctypes.c_void_p(bufs[r]), n, 2, _lib.ptr(outs[r]), st), "tq_p2p_wait_sum")
            torch.cuda.synchronize()
            expect = vals[0] + vals[1]                       # rank order: identical bits on both ranks
            assert torch.equal(outs[0], expect) and torch.equal(outs[1], expect), call
        seq = ctypes.c_uint64(1)
        _lib.check(lib.tq_p2p_timed_out(ctypes.c_void_p(bufs[0]), ctypes.byref(seq)), "tq_p2p_timed_out")
        assert seq.value == 0
    finally:
        for b in bufs:
            lib.tq_p2p_free(ctypes.c_void_p(b))


def test_p2p_rejects_oversized_requests():
    from tapqir_b200 import _lib

    lib = _lib.load()
    dummy = torch.zeros(4, dtype=torch.float64, device="cuda")
    peers = torch.zeros(2, dtype=torch.int64, device="cuda")
    with pytest.raises(ValueError):
        _lib.check(lib.tq_p2p_push(_lib.ptr(dummy), lib.tq_p2p_max_values() + 1, 0, 2, _lib.ptr(peers), _lib.stream_ptr(dummy.device)),
                   "tq_p2p_push")
