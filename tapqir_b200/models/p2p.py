"""
Host side of the NVLink peer-memory all-reduce of csrc/p2p_allreduce.cu: allocates this rank's buffer, exchanges the
CUDA IPC handles through the process group, opens the peers' buffers and keeps the device array of pointers the push
kernel needs.  One instance per engine; `push` / `wait_sum` enqueue one small kernel each (CUDA-graph capturable).
"""

import ctypes

import torch

from tapqir_b200 import _lib


class PeerMemoryUnavailable(RuntimeError):
    """Raised by EVERY rank when any rank could not allocate / map the peer buffers."""


class P2PAllReduce:
    def __init__(self, device, rank, world_size, group=None):
        lib = self.lib = _lib.load()
        if world_size > lib.tq_p2p_max_ranks():
            raise PeerMemoryUnavailable(f"at most {lib.tq_p2p_max_ranks()} ranks")
        self.device, self.rank, self.world = torch.device(device), int(rank), int(world_size)
        self._opened, self.own, failure = [], None, None
        with torch.cuda.device(self.device):
            # every rank goes through every collective below whatever happens to its own CUDA calls: a rank that failed
            # sends None / votes 0 instead of leaving the others waiting in a barrier
            own, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
            try:
                _lib.check(lib.tq_p2p_alloc(ctypes.byref(own), ctypes.cast(handle, ctypes.c_void_p)), "tq_p2p_alloc")
                self.own = own.value
            except Exception as err:
                failure = err
            handles = [None] * self.world
            torch.distributed.all_gather_object(handles, bytes(handle) if failure is None else None, group=group)
            ptrs = []
            if failure is None and all(h is not None for h in handles):
                try:
                    for q, h in enumerate(handles):
                        if q == self.rank:
                            ptrs.append(self.own)
                            continue
                        peer = ctypes.c_void_p()
                        raw = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                        _lib.check(lib.tq_p2p_open(ctypes.cast(raw, ctypes.c_void_p), ctypes.byref(peer)), "tq_p2p_open")
                        self._opened.append(peer.value)
                        ptrs.append(peer.value)
                except Exception as err:
                    failure = err
            elif failure is None:
                failure = RuntimeError("a peer could not allocate its buffer")
            backend = torch.distributed.get_backend(group)
            vote = torch.tensor([0 if failure is not None else 1], dtype=torch.int32,
                                device=self.device if backend == "nccl" else "cpu")
            torch.distributed.all_reduce(vote, op=torch.distributed.ReduceOp.MIN, group=group)
            if int(vote.item()) == 0:
                self.close()
                raise PeerMemoryUnavailable(str(failure) if failure is not None else "peer access failed on another rank")
            self.peers = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
            torch.cuda.synchronize(self.device)
        torch.distributed.barrier(group=group)   # every rank has every buffer mapped before anyone pushes

    def push(self, values, stream_ptr):
        n = values.numel()
        assert values.dtype == torch.float64 and n <= self.lib.tq_p2p_max_values()
        _lib.check(self.lib.tq_p2p_push(_lib.ptr(values), n, self.rank, self.world, _lib.ptr(self.peers), stream_ptr), "tq_p2p_push")

    def wait_sum(self, out, stream_ptr):
        _lib.check(self.lib.tq_p2p_wait_sum(ctypes.c_void_p(self.own), out.numel(), self.world, _lib.ptr(out), stream_ptr),
                   "tq_p2p_wait_sum")

    def timed_out(self):
        """Sequence number of the first call whose wait gave up (0 = never): a dead or stalled peer.  Synchronises."""
        seq = ctypes.c_uint64()
        _lib.check(self.lib.tq_p2p_timed_out(ctypes.c_void_p(self.own), ctypes.byref(seq)), "tq_p2p_timed_out")
        return seq.value

    def close(self):
        """Unmap the peers' buffers and free this rank's own (cudaMalloc'ed outside the caching allocator)."""
        for p in self._opened:
            self.lib.tq_p2p_close(ctypes.c_void_p(p))
        self._opened = []
        if self.own is not None:
            self.lib.tq_p2p_free(ctypes.c_void_p(self.own))
            self.own = None
