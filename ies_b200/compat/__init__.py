"""Stand-ins that let the reference's scripts run UNCHANGED on the b200 engine.

The scripts import `mpi4py`, `matplotlib`, `h5py` (and the engine modules as top-level
`space`, `source`, `structure`, `collector`, `plotter`, `recorder`) at module level
(tutorials/RT_simple_slabs.py:2-11, examples/mie/mie_scattering.py:1-11).  This image has
none of the three third-party packages; the engine needs none of them.

    python -m ies_b200.compat.run <script.py> <script args ...>

installs, for the MISSING packages only, minimal stand-ins in sys.modules, puts the engine's
directory on sys.path (so `import space` resolves to ies_b200/space.py, as the reference's
`sys.path.append(<library dir>)` expects) and executes the script text as `__main__`.

* mpi4py.MPI : COMM_WORLD = the engine's communicator (comm.default_comm(): one rank, or the
               CUDA-IPC communicator of a multi-process launch) with the calls the reference uses
               (SURVEY.md 2.2): Get_rank/Get_size/Barrier/barrier/gather; Get_processor_name().
* matplotlib : no-op figure API (plots are out of the hot-path scope); a real matplotlib is used
               when it is installed.
* h5py       : absent -> the engine's save_* functions write .npz with the same dataset names.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
ENGINE_DIR = os.path.dirname(_HERE)


def _missing(name):
    try:
        importlib.import_module(name)
        return False
    except Exception:
        return True


class _Anything:
    """Object that absorbs any attribute access / call / indexing (no-op plotting API)."""

    def __getattr__(self, k): return _Anything()
    def __call__(self, *a, **k): return _Anything()
    def __getitem__(self, k): return _Anything()
    def __iter__(self): return iter((_Anything(), _Anything()))
    def __enter__(self): return self
    def __exit__(self, *a): return False


class _NoopModule(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return _Anything()


def _install_matplotlib():
    names = ('matplotlib', 'matplotlib.pyplot', 'matplotlib.ticker', 'matplotlib.cm', 'matplotlib.colors',
             'mpl_toolkits', 'mpl_toolkits.mplot3d', 'mpl_toolkits.mplot3d.axes3d', 'mpl_toolkits.axes_grid1')
    for n in names:
        m = _NoopModule(n)
        m.__path__ = []
        sys.modules[n] = m
        if '.' in n:
            setattr(sys.modules[n.rsplit('.', 1)[0]], n.rsplit('.', 1)[1], m)
    sys.modules['matplotlib'].use = lambda *a, **k: None


def _install_mpi4py():
    if ENGINE_DIR not in sys.path:
        sys.path.insert(0, ENGINE_DIR)
    import comm as _comm      # the engine's communicator module (top-level, like `space`)

    class _MPI(types.ModuleType):
        @property
        def COMM_WORLD(self):
            return _comm.default_comm()

        @staticmethod
        def Get_processor_name():
            return os.uname().nodename

    pkg = types.ModuleType('mpi4py')
    pkg.__path__ = []
    mpi = _MPI('mpi4py.MPI')
    pkg.MPI = mpi
    sys.modules['mpi4py'] = pkg
    sys.modules['mpi4py.MPI'] = mpi


def install():
    """Idempotent.  Returns the list of packages that were replaced by stand-ins."""
    done = []
    if ENGINE_DIR not in sys.path:
        sys.path.insert(0, ENGINE_DIR)
    if _missing('mpi4py'):
        _install_mpi4py(); done.append('mpi4py')
    if _missing('matplotlib'):
        _install_matplotlib(); done.append('matplotlib')
    return done
