"""ctypes binding of libies_b200.so (include/ies_b200.h).

There is no CPU fallback: if the shared library is missing or cannot be
loaded, importing the engine raises, and every compute call needs a CUDA
device.  Build the library with `python -c "import __graft_entry__ as g; g.build()"`
or `ies_b200/csrc/build.sh`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('IES_B200_LIB') or os.path.join(_HERE, 'libies_b200.so')     # override: A/B of two builds

F32, F64, C64, C128 = 0, 1, 2, 3
FDTD, SHPF, PSTD = 0, 1, 2
HALF_H, HALF_E = 0, 1
COMP = {'Ex': 0, 'Ey': 1, 'Ez': 2, 'Hx': 3, 'Hy': 4, 'Hz': 5}
DTYPE_CODE = {np.dtype('float32'): F32, np.dtype('float64'): F64,
              np.dtype('complex64'): C64, np.dtype('complex128'): C128}
METHOD_CODE = {'FDTD': FDTD, 'SHPF': SHPF, 'PSTD': PSTD}
# derivative slots of one half-step (IES_D_*)
D_YFZ, D_ZFY, D_ZFX, D_XFZ, D_XFY, D_YFX = range(6)


class Config(C.Structure):
    _fields_ = [('nx', C.c_int32), ('ny', C.c_int32), ('nz', C.c_int32),
                ('dtype', C.c_int32), ('method', C.c_int32),
                ('rank', C.c_int32), ('nranks', C.c_int32), ('device', C.c_int32),
                ('dx', C.c_double), ('dy', C.c_double), ('dz', C.c_double), ('dt', C.c_double)]


class PmlTerm(C.Structure):
    _fields_ = [('half', C.c_int32), ('comp', C.c_int32), ('diff', C.c_int32), ('axis', C.c_int32),
                ('lo', C.c_int32 * 3), ('hi', C.c_int32 * 3),
                ('psi_off', C.c_int32), ('psi_thick', C.c_int32),
                ('sign', C.c_double),
                ('b', C.POINTER(C.c_double)), ('a', C.POINTER(C.c_double)), ('kf', C.POINTER(C.c_double))]


class EngineError(RuntimeError):
    pass


_lib = None
I3 = C.c_int32 * 3
I4 = C.c_int32 * 4
_vp = C.c_void_p
_dp = C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol declared in include/ies_b200.h
SIGNATURES = {
    'ies_last_error': (C.c_char_p, []),
    'ies_device_count': (C.c_int, [C.POINTER(C.c_int)]),
    'ies_create': (C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    'ies_destroy': (C.c_int, [_vp]),
    'ies_set_stream': (C.c_int, [_vp, _vp, C.c_int]),
    'ies_sync': (C.c_int, [_vp]),
    'ies_set_option': (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    'ies_set_coeff': (C.c_int, [_vp, C.c_int, _dp, C.c_int64]),
    'ies_set_update_box': (C.c_int, [_vp, C.c_int, I3, I3]),
    'ies_set_multiplier': (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int32]),
    'ies_clear_pml': (C.c_int, [_vp]),
    'ies_add_pml_term': (C.c_int, [_vp, C.POINTER(PmlTerm)]),
    'ies_set_ghost': (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]),
    'ies_set_neighbours': (C.c_int, [_vp, C.c_int, C.c_int]),
    'ies_update_h': (C.c_int, [_vp, C.c_int64]),
    'ies_update_e': (C.c_int, [_vp, C.c_int64]),
    'ies_update_phase': (C.c_int, [_vp, C.c_int, C.c_int]),
    'ies_halo_send_ptr': (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(C.c_int64)]),
    'ies_halo_recv_ptr': (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(C.c_int64)]),
    'ies_halo_copy': (C.c_int, [_vp, _vp, C.c_int]),
    'ies_halo_ipc_export': (C.c_int, [_vp, _vp]),
    'ies_halo_ipc_connect': (C.c_int, [_vp, C.c_int, _vp]),
    'ies_halo_push': (C.c_int, [_vp, C.c_int]),
    'ies_halo_wait': (C.c_int, [_vp, C.c_int]),
    'ies_put_src': (C.c_int, [_vp, C.c_int, I3, I3, C.c_double, C.c_double, C.c_int, _vp, _vp, _vp]),
    'ies_get_field': (C.c_int, [_vp, C.c_int, I3, I3, _vp]),
    'ies_set_field': (C.c_int, [_vp, C.c_int, I3, I3, _vp]),
    'ies_field_ptr': (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    'ies_dft_create': (C.c_int, [_vp, I3, I3, I4, _dp, C.c_int32, C.POINTER(_vp)]),
    'ies_dft_accumulate': (C.c_int, [_vp, _vp, _vp, C.c_int64]),
    'ies_dft_read': (C.c_int, [_vp, C.c_int, _vp]),
    'ies_dft_destroy': (C.c_int, [_vp]),
    'ies_probe_create': (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.POINTER(_vp)]),
    'ies_probe_record': (C.c_int, [_vp, _vp, _vp, C.c_int64]),
    'ies_probe_read': (C.c_int, [_vp, C.c_int, _vp]),
    'ies_probe_destroy': (C.c_int, [_vp]),
    'ies_launch_count': (C.c_int64, []),
    'ies_timer_start': (C.c_int, [_vp]),
    'ies_timer_stop': (C.c_int, [_vp, C.POINTER(C.c_double)]),
    'ies_profile': (C.c_int, [_vp, C.c_int]),
    'ies_fused_prof_read': (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    'ies_profile_read': (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
}


def load():
    """Load the shared library (once).  Raises EngineError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise EngineError(f"{LIB_PATH} not found: the CUDA engine is not built "
                          f"(run ies_b200/csrc/build.sh); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            if 'IES_B200_LIB' in os.environ:      # an older build under A/B comparison lacks newer entry points
                continue
            raise EngineError(f"{LIB_PATH} does not export {name}: stale build, run ies_b200/csrc/build.sh")
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise EngineError(load().ies_last_error().decode())


def dptr(a):
    return a.ctypes.data_as(_dp)
