"""GPU: error behaviour of the host mirror and of the C-ABI -- the same exception types at the
same points as the reference (space.py:51-52, 89, 99, 107-108, 162, 573-574; source.py:125-128,
222, 253), and loud failures (no silent other path) for what the engine does not cover."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
um = 1e-6
GAP = (720 * um / 32, 512 * um / 16, 512 * um / 16)
DT = 0.25 * min(GAP) / 299792458.0


def mk(product, grid=(32, 16, 16), dtype=np.float64, method='SHPF', gap=GAP, dt=DT, **kw):
    return product.space.Basic3D(grid, gap, dt, 10, dtype, np.complex128, method=method, engine='b200', **kw)


def test_constructor_asserts(product):
    with pytest.raises(AssertionError):
        mk(product, grid=(32, 16))                         # space.py:51
    with pytest.raises(AssertionError):
        mk(product, dt=DT * 100)                           # causality check, space.py:107-108
    with pytest.raises(ValueError):
        product.space.Basic3D((32, 16, 16), GAP, DT, 10, np.int32, np.complex128, method='SHPF', engine='b200')   # space.py:162
    with pytest.raises((AssertionError, NotImplementedError)):
        mk(product, method='HPF')                          # unfinished in the reference, not offered here


def test_spectral_axes_of_any_length_are_accepted(product):
    """The reference takes any N (space.py:145-162).  Powers of two in 16..512 run the FFT kernels, every
    other length the direct-circulant path (parity: the *_50cube / *_odd cases of test_gpu_parity.py); a
    real field dtype on an ODD axis is refused like the reference's irfftn would fail on it."""
    sp = mk(product, grid=(32, 20, 16), gap=(720 * um / 32, 512 * um / 20, 512 * um / 16))
    assert sp.Ex.shape == (32, 20, 16)
    sp = mk(product, grid=(30, 20, 18), method='FDTD', gap=(720 * um / 30, 512 * um / 20, 512 * um / 18),
            dt=0.25 * min(720 * um / 30, 512 * um / 20, 512 * um / 18) / 299792458.0)
    assert sp.Ex.shape == (30, 20, 18)                     # FDTD takes any grid
    sp = mk(product, grid=(32, 21, 16), gap=(720 * um / 32, 512 * um / 21, 512 * um / 16))
    sp.malloc()
    sp.apply_PML({'x': '+-', 'y': '', 'z': ''}, 4)
    sp.apply_BBC({'x': False, 'y': False, 'z': False}); sp.apply_PBC({'x': False, 'y': True, 'z': True})
    sp.init_update_constants()
    with pytest.raises(ValueError):
        sp.updateH(0)


def test_bloch_needs_complex_fields_and_a_setter(product):
    sp = mk(product)
    sp.malloc()
    sp.apply_PML({'x': '+-', 'y': '', 'z': ''}, 4)
    with pytest.raises(AssertionError):                    # space.py:573-574
        sp.apply_BBC({'x': False, 'y': True, 'z': True})


def test_update_before_init_constants_raises(product):
    sp = mk(product)
    sp.malloc()
    sp.apply_PML({'x': '+-', 'y': '', 'z': ''}, 4)
    sp.apply_BBC({'x': False, 'y': False, 'z': False})
    sp.apply_PBC({'x': False, 'y': True, 'z': True})
    with pytest.raises(RuntimeError):
        sp.updateH(0)


def test_setter_errors(product):
    sp = mk(product)
    sp.malloc()
    sp.apply_PML({'x': '+-', 'y': '', 'z': ''}, 4)
    sp.apply_BBC({'x': False, 'y': False, 'z': False})
    sp.apply_PBC({'x': False, 'y': True, 'z': True})
    with pytest.raises(ValueError):                        # source.py:125-128
        product.source.Setter(sp, (300 * um, 0, 0), (200 * um, 512 * um, 512 * um), (0, 0, 0))
    st = product.source.Setter(sp, (200 * um, 0, 0), (200 * um, 512 * um, 512 * um), (0, 0, 0))
    sp.init_update_constants()
    with pytest.raises(ValueError):                        # source.py:253
        st.put_src('Ey', 1.0, 'medium')
    with pytest.raises(TypeError):                         # complex pulse into real fields, source.py:222
        st.put_src('Ey', 1.0 + 1.0j, 'soft')


def test_c_abi_reports_errors_instead_of_falling_back(product):
    from ies_b200 import _lib
    lib = _lib.load()
    cfg = _lib.Config(8, 16, 16, 1, 1, 0, 1, 99, 1e-6, 1e-6, 1e-6, 1e-16)     # device 99 does not exist
    ctx = C.c_void_p()
    assert lib.ies_create(C.byref(cfg), C.byref(ctx)) != 0
    assert b'device' in lib.ies_last_error()
    cfg = _lib.Config(8, 1, 16, 1, 1, 0, 1, 0, 1e-6, 1e-6, 1e-6, 1e-16)       # a spectral axis of one point
    assert lib.ies_create(C.byref(cfg), C.byref(ctx)) != 0
    assert b'bad grid' in lib.ies_last_error()
    cfg = _lib.Config(8, 24, 16, 1, 1, 0, 1, 0, 1e-6, 1e-6, 1e-6, 1e-16)      # ny = 24 with SHPF: direct-circulant path
    assert lib.ies_create(C.byref(cfg), C.byref(ctx)) == 0
    assert lib.ies_destroy(ctx) == 0
    sp = mk(product)
    assert lib.ies_update_phase(sp._ctx, 5, 0) != 0                            # bad half
    assert lib.ies_set_option(sp._ctx, b'no_such_option', 1) != 0
    assert lib.ies_update_h(sp._ctx, 0) != 0                                   # no coefficients uploaded yet
    assert b'init_update_constants' in lib.ies_last_error()
