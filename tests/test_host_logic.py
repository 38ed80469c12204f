"""CPU: host-side logic of the product (no GPU needed) -- Setter / collector / structure
index arithmetic against the oracle and against rasters produced by the real reference,
the CPML term table, the spectral multiplier tables, and the N-rank exchange pattern over
gloo (world_size 2)."""
import os
import types

import numpy as np
import pytest
from scipy.constants import c, epsilon_0, mu_0

from oracle import ies_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class FakeComm:
    def __init__(self, rank, size): self.rank, self.size = rank, size
    def Get_rank(self): return self.rank
    def Get_size(self): return self.size
    def Barrier(self): pass
    def barrier(self): pass


def fake_space(grid, gap, rank=0, size=1, dtype=np.float64, tsteps=10):
    """A host-only stand-in exposing what Setter/collector/structure read from a space."""
    Nx, Ny, Nz = grid
    myNx = Nx // size
    sp = types.SimpleNamespace(
        Nx=Nx, Ny=Ny, Nz=Nz, dx=gap[0], dy=gap[1], dz=gap[2], grid=grid, dimension=3, tsteps=tsteps,
        MPIrank=rank, MPIsize=size, MPIcomm=FakeComm(rank, size), field_dtype=dtype, myNx=myNx,
        loc_grid=(myNx, Ny, Nz), myNx_indice=[(r * myNx, (r + 1) * myNx) for r in range(size)],
        BBC_called=False, _dirty=False, mmt=None)
    sp.eps = np.ones(sp.loc_grid) * epsilon_0
    sp.mu = np.ones(sp.loc_grid) * mu_0
    sp.eps_Ex = sp.eps_Ey = sp.eps_Ez = sp.eps
    sp.mu_Hx = sp.mu_Hy = sp.mu_Hz = sp.mu
    return sp


GRID, GAP = (48, 20, 18), (15e-6, 25.6e-6, 28.4e-6)


@pytest.mark.parametrize('size', [1, 2, 4, 8])
@pytest.mark.parametrize('xs', [0.0, 14.9e-6, 22.5e-6, 37.5e-6, 0.2 * 720e-6, 359.9e-6, 600e-6])
def test_setter_indices_match_oracle(size, xs):
    from ies_b200 import source
    for rank in range(size):
        sp = fake_space(GRID, GAP, rank, size)
        osp = O.OracleSpace(GRID, GAP, 1e-15, 10, np.float64, np.complex128, method='FDTD', rank=rank, size=size)
        for (s0, s1) in (((xs, 0, 0), (xs, 512e-6, 512e-6)),
                         ((xs, 100e-6, 200e-6), (xs + GAP[0], 100e-6 + GAP[1], 200e-6 + GAP[2]))):
            if round(s1[0] / GAP[0]) - 1 < 0 and round(s0[0] / GAP[0]) == round(s1[0] / GAP[0]):
                continue
            a = source.Setter(sp, s0, s1, (0., 1e4, 2e4))
            b = O.OracleSetter(osp, s0, s1, (0., 1e4, 2e4))
            assert (a.src_xsrt, a.src_xend, a.src_ysrt, a.src_yend, a.src_zsrt, a.src_zend, a.who_put_src) == \
                   (b.src_xsrt, b.src_xend, b.src_ysrt, b.src_yend, b.src_zsrt, b.src_zend, b.who_put_src)
            if a.who_put_src == rank:
                assert (a.my_src_xsrt, a.my_src_xend) == (b.my_src_xsrt, b.my_src_xend)
                assert np.array_equal(a.py, b.py) and np.array_equal(a.pz, b.pz) and np.array_equal(a.px, b.px)


@pytest.mark.parametrize('size', [1, 2, 3, 4])
def test_local_x_loc_matches_oracle(size):
    from ies_b200 import collector, structure
    for rank in range(size):
        sp = fake_space(GRID, GAP, rank, size)
        osp = O.OracleSpace(GRID, GAP, 1e-15, 10, np.float64, np.complex128, method='FDTD', rank=rank, size=size)
        col = collector.collector.__new__(collector.collector); col.space = sp
        st = structure.Structure('s', sp)
        for a in range(0, 47, 5):
            for b in range(a, 48, 7):
                want = O.local_x_loc(osp, a, b)
                assert col._get_local_x_loc(a, b) == want
                assert st._get_local_x_loc(a, b) == want


def test_structures_match_reference_rasters():
    """Box / Sphere / Cylinder3D rasters vs the real reference (tests/golden/structures.npz,
    written by oracle/make_structure_golden.py)."""
    from ies_b200 import structure
    z = np.load(os.path.join(GOLD, 'structures.npz'))
    grid, gap = tuple(int(v) for v in z['grid']), tuple(float(v) for v in z['gap'])
    for size in (1, 2):
        parts_e, parts_m = [], []
        for rank in range(size):
            sp = fake_space(grid, gap, rank, size)
            structure.Box('b', sp, (60e-6, 0, 0), (150e-6, 300e-6, 512e-6), 4., 1.)
            structure.Sphere('s', sp, (24, 10, 9), 120e-6, 2.25, 1.5)
            structure.Cylinder3D('c', sp, 'x', 90e-6, (400e-6, 560e-6), (256e-6, 256e-6), 6., 1.)
            structure.Cylinder3D('d', sp, 'y', 60e-6, (100e-6, 400e-6), (620e-6, 300e-6), 3., 2.)
            parts_e.append(sp.eps); parts_m.append(sp.mu)
        assert np.array_equal(np.concatenate(parts_e, 0), z['eps'])
        assert np.array_equal(np.concatenate(parts_m, 0), z['mu'])


def test_full_multiplier_is_hermitian_extension():
    """Real fields: the full-spectrum table reproduces irfftn(mult * rfftn(x)) exactly as a
    complex FFT of two packed lines."""
    from ies_b200.space import Basic3D
    n, d = 32, 1.3e-6
    k = np.fft.rfftfreq(n, d) * 2 * np.pi
    ik = (1j * k).astype(np.complex128)
    shift = np.exp(ik * d / 2)
    self = types.SimpleNamespace(field_dtype=np.float64, mmtdtype=np.complex128)
    full = Basic3D._full_multiplier(self, ik, shift, 0., n)
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    want_a = np.fft.irfft(ik * shift * np.fft.rfft(a), n)
    want_b = np.fft.irfft(ik * shift * np.fft.rfft(b), n)
    got = np.fft.ifft(full * np.fft.fft(a + 1j * b))
    assert np.abs(got.real - want_a).max() <= 1e-9 * np.abs(want_a).max()
    assert np.abs(got.imag - want_b).max() <= 1e-9 * np.abs(want_b).max()


@pytest.mark.parametrize('method', ['FDTD', 'SHPF', 'PSTD'])
@pytest.mark.parametrize('rank,size', [(0, 1), (0, 2), (1, 2), (1, 3)])
def test_pml_term_table_matches_oracle_slices(method, rank, size):
    """The product's CPML boxes equal the oracle's (reference-pinned) slice table."""
    if method == 'PSTD' and size > 1:
        pytest.skip('PSTD is single rank')
    from ies_b200.space import Basic3D
    grid, P = (24, 20, 16), 4
    gap = (1e-6, 1.1e-6, 1.2e-6)
    region = {'x': '+-', 'y': '+-', 'z': '+-'}
    osp = O.OracleSpace(grid, gap, 1e-16, 10, np.float64, np.complex128, method=method, rank=rank, size=size)
    osp.apply_PML(region, P)
    self = types.SimpleNamespace(method=method, MPIrank=rank, MPIsize=size, npml=P, PMLregion=region,
                                 loc_grid=osp.loc_grid, Ny=grid[1], Nz=grid[2])
    for ax in 'xyz':
        for nm in ('PMLb', 'PMLa', 'PMLkappa'):
            setattr(self, nm + ax, {'PMLb': osp.pml[ax]['b'], 'PMLa': osp.pml[ax]['a'],
                                    'PMLkappa': osp.pml[ax]['kappa']}[nm])
    for f in ('_main_boxes', '_pml_faces', '_pml_axis_rule', '_pml_terms'):
        setattr(self, f, types.MethodType(getattr(Basic3D, f), self))
    self._PML_ROWS = Basic3D._PML_ROWS
    terms = self._pml_terms()
    want = []
    for half in 'HE':
        for face in osp._pml_faces():
            prof, rows = osp._face_rows(half, face)
            for tgt, dn, sgn, fs, ps, psn in rows:
                box = [s.indices(n)[:2] for s, n in zip(fs, osp.loc_grid)]
                pbox = [s.indices(n)[:2] for s, n in zip(ps, osp.psi[f'{psn}_{face[1]}'].shape)]
                a = 'xyz'.index(face[0])
                want.append((half, tgt, tuple(b[0] for b in box), tuple(max(b) for b in box),
                             pbox[a][0] - 0, float(sgn), tuple(osp.pml[face[0]]['b'][prof])))
    got = [('HE'[t['half']], ('H' if t['half'] == 0 else 'E') + 'xyz'[t['comp']], tuple(t['lo']), tuple(t['hi']),
            t['psi_off'], t['sign'], tuple(t['b'])) for t in terms]
    assert got == want


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tests import torch_comm
    tc = torch_comm.TorchComm()
    ok = True
    for step in range(3):
        first = [torch.full((4, 5), float(100 * rank + step)), torch.full((4, 5), float(100 * rank + step + 0.5))]
        last = [torch.full((4, 5), float(-100 * rank - step)), torch.full((4, 5), float(-100 * rank - step - 0.5))]
        recv_h = [torch.zeros(4, 5), torch.zeros(4, 5)]
        recv_e = [torch.zeros(4, 5), torch.zeros(4, 5)]
        tc.exchange_planes(0, first, recv_h)       # updateH: my first planes -> rank-1, recv rank+1's
        tc.exchange_planes(1, last, recv_e)        # updateE: my last planes -> rank+1, recv rank-1's
        if rank < world - 1:
            ok &= bool(recv_h[0][0, 0] == 100 * (rank + 1) + step and recv_h[1][0, 0] == 100 * (rank + 1) + step + 0.5)
        else:
            ok &= bool(recv_h[0].abs().sum() == 0)
        if rank > 0:
            ok &= bool(recv_e[0][0, 0] == -100 * (rank - 1) - step and recv_e[1][0, 0] == -100 * (rank - 1) - step - 0.5)
        else:
            ok &= bool(recv_e[0].abs().sum() == 0)
    g = tc.gather(np.full(3, rank), root=0)
    if rank == 0:
        ok &= [int(a[0]) for a in g] == list(range(world))
    tc.Barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


@pytest.mark.parametrize('world', [2, 3])
def test_torchcomm_exchange_pattern_gloo(world):
    """N > 1 plumbing on CPU: the neighbour send/recv pattern (first planes to rank-1 before
    updateH, last planes to rank+1 before updateE) over gloo, tests/torch_comm.TorchComm."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    ps = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = [q.get(timeout=180) for _ in ps]
    [p.join(60) for p in ps]
    assert all(ok for _, ok in res), res


class _FakeHaloLib:
    """CPU stand-in for the four ies_halo_* entry points IpcComm drives: a 'handle' is the rank
    number, a push appends (half, seq) to a file of the neighbour it targets, a wait records what
    the rank expects.  Lets the connect / push / wait call pattern run without a GPU."""

    def __init__(self, rank, size, tmp):
        self.rank, self.size, self.tmp = rank, size, tmp
        self.peers, self.seq, self.waits = {}, [0, 0], []

    def ies_halo_ipc_export(self, ctx, h):
        h[0] = self.rank + 1
        return 0

    def ies_halo_ipc_connect(self, ctx, nbr, h):
        self.peers[nbr] = int(h[0]) - 1
        return 0

    def ies_halo_push(self, ctx, half):
        nbr = 0 if half == 0 else 1
        if nbr in self.peers:
            self.seq[half] += 1
            with open(os.path.join(self.tmp, f'to{self.peers[nbr]}_half{half}'), 'a') as f:
                f.write(f'{self.rank} {self.seq[half]}\n')
        return 0

    def ies_halo_wait(self, ctx, half):
        src = self.rank + 1 if half == 0 else self.rank - 1
        if 0 <= src < self.size:
            self.waits.append((half, src))
        return 0

    def ies_last_error(self): return b''


def _ipc_worker(rank, world, port, tmp, q):
    import types
    from ies_b200 import comm, _lib
    fake = _FakeHaloLib(rank, world, tmp)
    _lib._lib = fake                                   # IpcComm calls _lib.load()
    c = comm.IpcComm(rank, world, comm.SocketStore(rank, world, port=port))
    spaces = [types.SimpleNamespace(_ctx=None), types.SimpleNamespace(_ctx=None)]     # TF and IF
    for step in range(3):
        for sp in spaces: c.exchange(sp, 0)
        for sp in spaces: c.exchange(sp, 1)
    c.Barrier()
    g = c.gather(np.full(3, rank), root=0)
    ok = fake.peers == {n: r for n, r in ((0, rank - 1), (1, rank + 1)) if 0 <= r < world}
    exp_waits = [(h, s) for h, s in ((0, rank + 1), (1, rank - 1)) if 0 <= s < world]
    ok &= fake.waits == [w for _ in range(3) for w in ([x for x in exp_waits if x[0] == 0] * 2 + [x for x in exp_waits if x[0] == 1] * 2)]
    if rank == 0:
        ok &= [int(a[0]) for a in g] == list(range(world))
    c.Barrier()
    # what arrived for me: half 0 pushes come from rank+1, half 1 pushes from rank-1, 6 each (2 spaces x 3 steps)
    for half, src in ((0, rank + 1), (1, rank - 1)):
        path = os.path.join(tmp, f'to{rank}_half{half}')
        if 0 <= src < world:
            rows = [l.split() for l in open(path)]
            ok &= len(rows) == 6 and all(int(r[0]) == src for r in rows)
        else:
            ok &= not os.path.exists(path)
    q.put((rank, bool(ok)))


@pytest.mark.parametrize('world', [2, 3])
def test_ipccomm_rendezvous_and_exchange_pattern(world, tmp_path):
    """N > 1 plumbing on CPU: SocketStore rendezvous (set/get/barrier/gather over TCP) and the
    export -> connect -> push -> wait sequence IpcComm drives, against a fake halo library."""
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 30600 + world + (os.getpid() % 200)
    ps = [ctx.Process(target=_ipc_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(60) for p in ps]
    assert all(ok for _, ok in res), res


@pytest.mark.parametrize('n', [20, 50, 36])
def test_full_multiplier_on_even_non_power_of_two_axes(n):
    """Axes that are not a power of two take the engine's direct-circulant path; the host table is the
    same Hermitian extension, and ifft(full * fft(x)) must equal the reference's irfft(m * rfft(x))."""
    import types
    from ies_b200.space import Basic3D
    rng = np.random.default_rng(n)
    d = 3.1e-6
    k = np.fft.rfftfreq(n, d) * 2 * np.pi
    ik = (1j * k).astype(np.complex128)
    shift = np.exp(ik * d / 2)
    self = types.SimpleNamespace(field_dtype=np.float64, mmtdtype=np.complex128)
    full = Basic3D._full_multiplier(self, ik, shift, 0., n)
    x = rng.standard_normal(n)
    want = np.fft.irfft(ik * shift * np.fft.rfft(x), n)
    got = np.fft.ifft(full * np.fft.fft(x))
    assert np.max(np.abs(got.imag)) < 1e-9 * np.max(np.abs(want))
    assert np.allclose(got.real, want, rtol=0, atol=1e-12 * np.max(np.abs(want)))
    with pytest.raises(ValueError):
        Basic3D._full_multiplier(self, ik, shift, 0., n + 1)
