"""torch.distributed neighbour exchange (NCCL / gloo) -- TEST INFRASTRUCTURE.

Round 1's transport of the slab halo planes; the product now moves them itself (comm.IpcComm:
CUDA IPC + copy engine).  Kept for the world_size-2 gloo test of the exchange pattern and as a
cross-check transport on GPU boxes."""
import os

from ies_b200 import _lib


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


