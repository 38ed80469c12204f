"""x-slab communicators for the b200 engine.

The reference hands one yz-plane to a neighbour rank per half-step with blocking
pickled mpi4py send/recv (space.py:645-670, 863-887) and uses Barrier/gather
elsewhere.  Here a communicator object offers the same rank/size
view plus `exchange(space, half)`:

* SingleComm  -- one slab, nothing to exchange.
* LocalComm   -- several slabs driven by ONE process (same or different GPUs):
                 cudaMemcpyPeerAsync between contexts, event ordered (ies_halo_copy).
* TorchComm   -- one process per GPU (torchrun): torch.distributed point-to-point
                 (NCCL over NVLink for CUDA planes; gloo for the CPU tests).  The
                 planes are the engine's own device buffers wrapped as tensors --
                 torch is plumbing only.
"""
import os

import numpy as np

try:
    from . import _lib
except ImportError:  # imported as a top-level module (sys.path drop-in)
    import _lib


class SingleComm:
    rank, size = 0, 1

    def Get_rank(self): return 0
    def Get_size(self): return 1
    def Barrier(self): pass
    def barrier(self): pass
    def exchange(self, space, half): pass
    def gather(self, arr, root=0): return [arr]


class LocalGroup:
    """All slabs of one decomposition inside this process."""

    def __init__(self, size, devices=None):
        self.size = size
        self.devices = list(devices) if devices is not None else [0] * size
        self.spaces = {}       # (space-group key, rank) -> space
        self.comms = [LocalComm(self, r) for r in range(size)]

    def comm(self, rank):
        return self.comms[rank]


class LocalComm:
    def __init__(self, group, rank):
        self.group, self.rank, self.size = group, rank, group.size
        self.device = group.devices[rank]
        self.members = None    # list of spaces of this decomposition, set by register()

    def Get_rank(self): return self.rank
    def Get_size(self): return self.size
    def Barrier(self): pass
    def barrier(self): pass

    def register(self, space, key):
        self.group.spaces[(key, self.rank)] = space
        space._local_key = key

    def exchange(self, space, half):
        """Pull the neighbour's planes into my halo buffers."""
        lib = _lib.load()
        src_rank = self.rank + 1 if half == _lib.HALF_H else self.rank - 1
        if src_rank < 0 or src_rank >= self.size:
            return
        src = self.group.spaces[(space._local_key, src_rank)]
        _lib.check(lib.ies_halo_copy(space._ctx, src._ctx, half))

    def gather(self, arr, root=0):
        raise NotImplementedError("LocalComm.gather: use ies_b200.space.gather_local")


class _DevPlane:
    """__cuda_array_interface__ view of an engine-owned device buffer."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {
            'shape': (nbytes,), 'typestr': '|u1', 'data': (int(ptr), False), 'version': 2, 'strides': None}


class TorchComm:
    """torch.distributed-backed neighbour exchange (one process per GPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self._views = {}
        self._streams = {}

    def Get_rank(self): return self.rank
    def Get_size(self): return self.size
    def Barrier(self): self.dist.barrier(self.group)
    def barrier(self): self.dist.barrier(self.group)

    # -- pattern shared with the gloo CPU tests ---------------------------------
    def exchange_planes(self, half, send, recv):
        """updateH (half 0): send my first planes to rank-1, receive rank+1's;
        updateE (half 1): send my last planes to rank+1, receive rank-1's.
        `send`/`recv` are lists of tensors (any device the backend supports)."""
        dist = self.dist
        dst = self.rank - 1 if half == 0 else self.rank + 1
        src = self.rank + 1 if half == 0 else self.rank - 1
        ops = []
        if 0 <= dst < self.size:
            ops += [dist.P2POp(dist.isend, t, dst, self.group) for t in send]
        if 0 <= src < self.size:
            ops += [dist.P2POp(dist.irecv, t, src, self.group) for t in recv]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def _tensors(self, space):
        import ctypes as C
        import torch
        # cached on the space itself (an id()-keyed dict would hand a new space the views of a
        # freed one that happened to get the same id)
        if getattr(space, '_halo_views', None) is None:
            lib = _lib.load()
            dev = torch.device('cuda', space.device)
            v = {}
            for half in (0, 1):
                for kind, fn in (('send', lib.ies_halo_send_ptr), ('recv', lib.ies_halo_recv_ptr)):
                    ts = []
                    for w in (0, 1):
                        p, n = C.c_void_p(), C.c_int64()
                        _lib.check(fn(space._ctx, half, w, C.byref(p), C.byref(n)))
                        ts.append(torch.as_tensor(_DevPlane(p.value, n.value), device=dev))
                    v[(half, kind)] = ts
            space._halo_views = v
        return space._halo_views

    def _stream_for(self, device):
        """One dedicated (non-default) torch stream per device, shared by NCCL and the engine.
        torch's default stream cannot be used: its handle is 0, which the C-ABI
        (ies_set_stream) reads as "use the context's own stream" -- the kernels would then
        run unordered against the NCCL transfers."""
        import torch
        if device not in self._streams:
            dev = device[0] if isinstance(device, tuple) else device
            self._streams[device] = torch.cuda.Stream(device=dev)
        return self._streams[device]

    def exchange(self, space, half):
        """In-order variant: transfer and kernels on one stream."""
        import torch
        v = self._tensors(space)
        st = self._stream_for(space.device)
        space._use_stream(st.cuda_stream)          # engine kernels and NCCL on the same stream
        with torch.cuda.stream(st):
            self.exchange_planes(half, v[(half, 'send')], v[(half, 'recv')])

    def exchange_begin(self, space, half):
        """Overlapped variant: the NCCL send/recv runs on a second stream once the engine
        stream has finished the previous update (event), and returns the event the
        neighbour-dependent part of the half-step has to wait for."""
        import torch
        v = self._tensors(space)
        se = self._stream_for(space.device)
        sc = self._stream_for((space.device, 'comm'))
        space._use_stream(se.cuda_stream)
        ready = torch.cuda.Event()
        ready.record(se)                           # fields of the previous half-step are final
        sc.wait_event(ready)
        with torch.cuda.stream(sc):
            self.exchange_planes(half, v[(half, 'send')], v[(half, 'recv')])
            done = torch.cuda.Event()
            done.record(sc)
        return done

    def exchange_end(self, space, done):
        self._stream_for(space.device).wait_event(done)

    def gather(self, arr, root=0):
        out = [None] * self.size if self.rank == root else None
        self.dist.gather_object(arr, out, dst=root, group=self.group)
        return out


def default_comm():
    """TorchComm when torch.distributed is initialised (torchrun), else SingleComm."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return TorchComm()
    except ImportError:
        pass
    return SingleComm()


def default_device(comm):
    if isinstance(comm, LocalComm):
        return comm.device
    if 'LOCAL_RANK' in os.environ and comm.size > 1:
        return int(os.environ['LOCAL_RANK'])
    return int(os.environ.get('IES_B200_DEVICE', 0))
