"""x-slab communicators for the b200 engine.

The reference hands one yz-plane to a neighbour rank per half-step with blocking
pickled mpi4py send/recv (space.py:645-670, 863-887) and uses Barrier/gather
elsewhere.  Here a communicator object offers the same rank/size
view plus `exchange(space, half)`:

* SingleComm  -- one slab, nothing to exchange.
* LocalComm   -- several slabs driven by ONE process (same or different GPUs):
                 cudaMemcpyPeerAsync between contexts, event ordered (ies_halo_copy).
* IpcComm     -- one process per GPU (torchrun / any launcher setting RANK, WORLD_SIZE,
                 LOCAL_RANK): CUDA IPC peer-mapped receive planes, copy-engine pushes over
                 NVLink, stream-ordered flags (ies_halo_push / ies_halo_wait); the ranks
                 find each other through SocketStore (standard library only -- no torch,
                 no MPI in the package).
"""
import os

import numpy as np

try:
    from . import _lib
except ImportError:  # imported as a top-level module (sys.path drop-in)
    import _lib


def local_x_loc(space, gxsrts, gxends):
    """Clip a global x range [gxsrts, gxends) to this rank's slab: (global range, local range) or
    (None, None).  ONE statement of the rule of structure.py:17-107 and collector.py:30-118 (the
    reference carries two copies): interval-overlap cases written as one intersection."""
    assert gxsrts >= 0
    assert gxends < space.Nx
    bxsrt, bxend = space.myNx_indice[space.MPIrank]
    lo, hi = max(gxsrts, bxsrt), min(gxends, bxend)
    # a zero-length request (gxsrts == gxends) belongs to the slab that contains gxsrts
    if (gxsrts == gxends and bxsrt <= gxsrts < bxend) or lo < hi:
        hi = max(hi, lo)
        return (lo, hi), (lo - bxsrt, hi - bxsrt)
    return None, None


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


class SocketStore:
    """Rendezvous for one-process-per-GPU runs on one node (the role `mpirun`'s wire-up plays
    for the reference, README.md:131-135): rank 0 serves a tiny key/value + barrier protocol
    over TCP on 127.0.0.1, the other ranks connect.  Pure standard library; values are bytes."""

    def __init__(self, rank, size, addr='127.0.0.1', port=None, timeout=300.):
        import socket
        import struct
        import threading
        self.rank, self.size, self.timeout = rank, size, timeout
        self._struct = struct
        if port is None:
            port = int(os.environ.get('IES_B200_PORT', int(os.environ.get('MASTER_PORT', '29500')) + 1017))
        self._server = None
        if rank == 0:
            self._kv, self._bar = {}, {}
            self._cv = threading.Condition()
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, port))
            srv.listen(size + 4)
            self._server = srv
            threading.Thread(target=self._accept_loop, daemon=True).start()
        import time
        t0 = time.time()
        while True:
            try:
                self._sock = socket.create_connection((addr, port), timeout=timeout)
                break
            except OSError:
                if time.time() - t0 > timeout:
                    raise
                time.sleep(0.05)
        self._sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        self._gen = 0

    # ---- wire format: 4-byte length + pickled tuple
    def _send(self, sock, obj):
        import pickle
        b = pickle.dumps(obj, protocol=4)
        sock.sendall(self._struct.pack('<I', len(b)) + b)

    def _recv(self, sock):
        import pickle

        def rd(n):
            buf = b''
            while len(buf) < n:
                c = sock.recv(n - len(buf))
                if not c:
                    raise ConnectionError('store connection closed')
                buf += c
            return buf
        (n,) = self._struct.unpack('<I', rd(4))
        return pickle.loads(rd(n))

    def _accept_loop(self):
        import threading
        while True:
            try:
                conn, _ = self._server.accept()
            except OSError:
                return
            threading.Thread(target=self._serve, args=(conn,), daemon=True).start()

    def _serve(self, conn):
        try:
            while True:
                msg = self._recv(conn)
                op = msg[0]
                if op == 'set':
                    with self._cv:
                        self._kv[msg[1]] = msg[2]
                        self._cv.notify_all()
                    self._send(conn, True)
                elif op == 'get':
                    with self._cv:
                        ok = self._cv.wait_for(lambda: msg[1] in self._kv, timeout=self.timeout)
                        val = self._kv.get(msg[1]) if ok else None
                    self._send(conn, val)
                elif op == 'del':
                    with self._cv:
                        for k in [k for k in self._kv if k.startswith(msg[1])]:
                            del self._kv[k]
                    self._send(conn, True)
                elif op == 'barrier':
                    with self._cv:
                        self._bar[msg[1]] = self._bar.get(msg[1], 0) + 1
                        self._cv.notify_all()
                        ok = self._cv.wait_for(lambda: self._bar[msg[1]] >= self.size, timeout=self.timeout)
                    self._send(conn, ok)
        except (ConnectionError, OSError, EOFError):
            return

    def _call(self, *msg):
        self._send(self._sock, msg)
        return self._recv(self._sock)

    def set(self, key, value): self._call('set', key, bytes(value))

    def get(self, key):
        v = self._call('get', key)
        if v is None:
            raise TimeoutError(f'store: key {key!r} never arrived')
        return v

    def delete_prefix(self, prefix): self._call('del', prefix)

    def barrier(self):
        self._gen += 1
        if not self._call('barrier', self._gen):
            raise TimeoutError('store: barrier timed out')


class IpcComm:
    """One process per GPU on one node: the slabs' halo planes travel by copy engine into
    peer-mapped (CUDA IPC) receive buffers, signalled by a stream-ordered flag write and
    consumed behind a stream memory wait -- ies_halo_push / ies_halo_wait of the C-ABI.
    Replaces the reference's blocking pickled mpi4py send/recv (space.py:645-670, 863-887);
    rank/size/Barrier/gather keep the mpi4py spelling the scripts use."""

    def __init__(self, rank=None, size=None, store=None):
        self.rank = int(os.environ['RANK']) if rank is None else rank
        self.size = int(os.environ['WORLD_SIZE']) if size is None else size
        self.store = store if store is not None else SocketStore(self.rank, self.size)
        self._nspaces = 0

    def Get_rank(self): return self.rank
    def Get_size(self): return self.size
    def Barrier(self): self.store.barrier()
    def barrier(self): self.store.barrier()

    def connect(self, space):
        """Publish this slab's IPC handle, map the neighbours'.  Spaces are created in the same
        order on every rank (SPMD script), so the k-th space of each rank forms one decomposition."""
        import ctypes as C
        lib = _lib.load()
        key = f'space{self._nspaces}'
        self._nspaces += 1
        h = (C.c_ubyte * 64)()
        _lib.check(lib.ies_halo_ipc_export(space._ctx, h))
        self.store.set(f'{key}/handle/{self.rank}', bytes(h))
        for nbr, r in ((0, self.rank - 1), (1, self.rank + 1)):
            if 0 <= r < self.size:
                hb = (C.c_ubyte * 64).from_buffer_copy(self.store.get(f'{key}/handle/{r}'))
                _lib.check(lib.ies_halo_ipc_connect(space._ctx, nbr, hb))
        self.store.barrier()          # everybody mapped everybody before the first push
        space._ipc_connected = True

    def exchange(self, space, half):
        """space.py:645-670 / 863-887: push my planes to the neighbour that needs them, then put the
        wait for the planes I need in front of the update."""
        if not getattr(space, '_ipc_connected', False):
            self.connect(space)
        lib = _lib.load()
        _lib.check(lib.ies_halo_push(space._ctx, half))
        _lib.check(lib.ies_halo_wait(space._ctx, half))

    def gather(self, arr, root=0):
        """mpi4py-style gather of picklable objects to `root` (plotter.Graphtool.gather)."""
        import pickle
        self._gat = getattr(self, '_gat', 0) + 1
        pre = f'gather{self._gat}/'
        self.store.set(pre + str(self.rank), pickle.dumps(arr, protocol=4))
        out = None
        if self.rank == root:
            out = [pickle.loads(self.store.get(pre + str(r))) for r in range(self.size)]
        self.store.barrier()
        if self.rank == root:
            self.store.delete_prefix(pre)
        return out


_default = None


def default_comm():
    """IpcComm when the process was started as one rank of several (RANK / WORLD_SIZE in the
    environment: torchrun, or any launcher that sets them), else SingleComm."""
    global _default
    if int(os.environ.get('WORLD_SIZE', '1')) > 1 and 'RANK' in os.environ:
        if _default is None:
            _default = IpcComm()
        return _default
    return SingleComm()


def default_device(comm):
    if isinstance(comm, LocalComm):
        return comm.device
    if 'LOCAL_RANK' in os.environ and comm.size > 1:
        return int(os.environ['LOCAL_RANK'])
    return int(os.environ.get('IES_B200_DEVICE', 0))
