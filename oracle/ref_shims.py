"""Load the UNMODIFIED reference modules from /root/reference on CPU -- test
infrastructure, build container only (the GPU box has no /root/reference).

The reference imports cupy, mpi4py, h5py and matplotlib at module top level;
none is installed here.  We inject stand-ins before loading:

* `cupy`      -> a module proxying NumPy (+ asnumpy/asarray), so the reference's
                 engine='cupy' branch (the only correct SHPF/PSTD branch at this
                 commit) runs the identical arithmetic on the CPU;
* `mpi4py.MPI`-> single-rank stub, or a threaded fake (`FakeComm`) implementing the
                 six calls the reference uses for N-rank runs;
* `h5py`, `matplotlib` -> empty stand-ins.

The y-/z+-/z- CPML functions index with Python *lists* of slices, an IndexError
on NumPy >= 1.23.  `load_reference()` patches `= [ ... slice ...]`
literals to tuples IN MEMORY while compiling space.py; nothing is written to
/root/reference and no reference source enters this repository.
"""
from __future__ import annotations

import importlib.util
import os
import queue
import re
import sys
import threading
import types

import numpy as np

def _ref_dir():
    # the reference where it lies in the build container, else the copy oracle/make_ref.py
    # placed under oracle/_ref/ (travels to the GPU box; never committed)
    for d in (os.environ.get('IES_REFERENCE_DIR'), '/root/reference',
              os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')):
        if d and os.path.isfile(os.path.join(d, 'space.py')):
            return d
    return '/root/reference'


REF = _ref_dir()


def reference_available():
    return os.path.isfile(os.path.join(REF, 'space.py'))


class _SingleComm:
    def Get_rank(self): return 0
    def Get_size(self): return 1
    def Barrier(self): pass
    def barrier(self): pass
    def gather(self, obj, root=0): return [obj]


class FakeWorld:
    """Threaded fake MPI world: one FakeComm per rank (thread)."""

    def __init__(self, size):
        self.size = size
        self.boxes = {}
        self.lock = threading.Lock()
        self.bar = threading.Barrier(size)

    def box(self, src, dst, tag):
        with self.lock:
            return self.boxes.setdefault((src, dst, tag), queue.Queue())


class FakeComm:
    def __init__(self, world, rank):
        self.world, self.rank = world, rank

    def Get_rank(self): return self.rank
    def Get_size(self): return self.world.size
    def Barrier(self): self.world.bar.wait()
    def barrier(self): self.world.bar.wait()

    def send(self, obj, dest, tag=0):
        self.world.box(self.rank, dest, tag).put(np.array(obj, copy=True))

    def recv(self, source, tag=0):
        return self.world.box(source, self.rank, tag).get(timeout=120)


class _Noop:
    """Absorbs any attribute access / call / indexing: the plotting calls of the reference's scripts."""
    def __getattr__(self, k): return _Noop()
    def __call__(self, *a, **k): return _Noop()
    def __getitem__(self, k): return _Noop()
    def __iter__(self): return iter((_Noop(), _Noop()))


class _NoopModule(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return _Noop()


_tls = threading.local()


class _MPIModule(types.ModuleType):
    @property
    def COMM_WORLD(self):
        return getattr(_tls, 'comm', None) or _SingleComm()

    @staticmethod
    def Get_processor_name():
        return 'oracle-host'


def set_thread_comm(comm):
    _tls.comm = comm


def _install_shims():
    if 'cupy' not in sys.modules or not getattr(sys.modules['cupy'], '_ies_shim', False):
        cp = types.ModuleType('cupy')
        cp.__dict__.update({k: getattr(np, k) for k in dir(np) if not k.startswith('__')})
        cp.asnumpy = np.asarray
        cp.asarray = np.asarray
        cp._ies_shim = True
        sys.modules['cupy'] = cp
    if 'mpi4py' not in sys.modules or not getattr(sys.modules['mpi4py'], '_ies_shim', False):
        pkg = types.ModuleType('mpi4py')
        pkg._ies_shim = True
        mpi = _MPIModule('mpi4py.MPI')
        pkg.MPI = mpi
        sys.modules['mpi4py'] = pkg
        sys.modules['mpi4py.MPI'] = mpi
    for name in ('h5py', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.ticker', 'mpl_toolkits',
                 'mpl_toolkits.mplot3d', 'mpl_toolkits.mplot3d.axes3d', 'mpl_toolkits.axes_grid1'):
        try:
            importlib.import_module(name)
        except Exception:
            mod = _NoopModule(name) if name != 'h5py' else types.ModuleType(name)
            mod.__path__ = []
            sys.modules[name] = mod
            if '.' in name:
                parent, child = name.rsplit('.', 1)
                setattr(sys.modules[parent], child, mod)
    ax = sys.modules['mpl_toolkits.axes_grid1']
    if not hasattr(ax, 'make_axes_locatable'):
        ax.make_axes_locatable = lambda *a, **k: None
    m3 = sys.modules['mpl_toolkits.mplot3d']
    if not hasattr(m3, 'axes3d'):
        m3.axes3d = sys.modules.get('mpl_toolkits.mplot3d.axes3d')


_LIST_LIT = re.compile(r'=\s*\[((?:\s*(?:None|slice\([^\)]*\))\s*,?)+)\]')


def _patch_list_indices(src):
    """`idx = [slice(..), None, ..]` -> `idx = (slice(..), None, ..)`."""
    return _LIST_LIT.sub(lambda m: '= (' + m.group(1) + ')', src)


_cache = {}


def load_reference(extra=()):
    """Returns a namespace with the reference's space/source/collector/structure modules
    (+ `extra`, e.g. plotter / recorder for the script-level goldens)."""
    if 'ns' in _cache and all(hasattr(_cache['ns'], n) for n in extra):
        return _cache['ns']
    if not reference_available():
        raise RuntimeError('the reference is not present on this machine (/root/reference or oracle/_ref)')
    _install_shims()
    ns = _cache.get('ns') or types.SimpleNamespace()
    for name in ('space', 'source', 'collector', 'structure') + tuple(extra):
        if hasattr(ns, name):
            continue
        path = os.path.join(REF, name + '.py')
        with open(path) as f:
            src = f.read()
        if name == 'space':
            src = _patch_list_indices(src)
        mod = types.ModuleType('ies_reference_' + name)
        mod.__file__ = path
        # the reference's modules import each other by their top-level names (recorder.py:2)
        saved = {n: sys.modules.get(n) for n in ('space', 'source', 'collector', 'structure')}
        for n in saved:
            if hasattr(ns, n):
                sys.modules[n] = getattr(ns, n)
        try:
            exec(compile(src, path, 'exec'), mod.__dict__)
        finally:
            for n, m in saved.items():
                if m is None:
                    sys.modules.pop(n, None)
                else:
                    sys.modules[n] = m
        setattr(ns, name, mod)
    _cache['ns'] = ns
    return ns


def run_ranks(size, fn):
    """Run fn(rank) on `size` threads, each seeing its own fake COMM_WORLD."""
    world = FakeWorld(size)
    out = [None] * size
    err = []

    def work(r):
        set_thread_comm(FakeComm(world, r))
        try:
            out[r] = fn(r)
        except BaseException as e:  # noqa
            err.append(e)
            try:
                world.bar.abort()
            except Exception:
                pass

    th = [threading.Thread(target=work, args=(r,)) for r in range(size)]
    [t.start() for t in th]
    [t.join() for t in th]
    if err:
        raise err[0]
    return out
