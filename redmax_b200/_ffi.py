"""ctypes binding of the C ABI in include/redmax_b200.h (the same symbols a MEX gateway binds).

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a) as redmax_b200/lib/libredmax_b200.so.
There is no CPU fallback: if the library is missing, or no CUDA device is present when a compute entry point is
called, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RMX_LIB: developer override to A/B a tagged build of the same CUDA library (tools/build_variant.sh); never a CPU path
LIB_PATH = os.environ.get('RMX_LIB') or os.path.join(_HERE, 'lib', 'libredmax_b200.so')

RMX_OK = 0
RMX_JOINT_FIXED = 0
RMX_JOINT_REVOLUTE = 1
RMX_JOINT_PRISMATIC = 2
RMX_JOINT_PLANAR = 3
RMX_JOINT_TRANSLATIONAL = 4
RMX_JOINT_FREE2D = 5
RMX_JOINT_UNIVERSAL = 6
RMX_JOINT_SPHERICAL = 7
RMX_JOINT_FREE3D = 8
RMX_MAX_JOINT_DOF = 6
RMX_FORCE_POINTPOINT = 0
RMX_FORCE_SPRINGDAMPER = 1
RMX_MAX_CABLE_POINTS = 4
RMX_SCHEME_BDF1 = 1
RMX_SCHEME_BDF2 = 2
RMX_LINSOLVE_LU = 0
RMX_LINSOLVE_PCG = 1
RMX_TAU_NONE = 0
RMX_TAU_CONST = 1
RMX_TAU_PER_STEP = 2
RMX_ST_DIVERGED = 1
RMX_ST_MAXITER = 2
RMX_ST_LSFAIL = 4
RMX_ST_NAN = 8
RMX_ST_SCHED = 16  # internal scheduling failure: the rollout's trajectory is not valid (Scene.rollout* raise on it)
RMX_ST_CHART = 32

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class rmx_scene_desc(C.Structure):
    _fields_ = [
        ('n', C.c_int32),
        ('parent', _pi), ('jtype', _pi),
        ('E0_pj', _pd), ('E0_ji', _pd), ('axis', _pd), ('axis2', _pd), ('I_i', _pd), ('sides', _pd),
        ('stiffness', _pd), ('damping', _pd), ('qRest', _pd),
        ('qLimL', _pd), ('qLimU', _pd), ('qLimK', _pd), ('qLimD', _pd),
        ('grav', C.c_double * 3),
        ('nground', C.c_int32),
        ('ground_body', _pi), ('ground_E', _pd),
        ('ground_kn', _pd), ('ground_kt', _pd), ('ground_kd', _pd), ('ground_mu', _pd),
        ('npointforce', C.c_int32),
        ('pf_body1', _pi), ('pf_body2', _pi), ('pf_x1', _pd), ('pf_x2', _pd), ('pf_ks', _pd), ('pf_kd', _pd),
        ('pf_kind', _pi), ('pf_L', _pd),
        ('ncable', C.c_int32),
        ('cable_npts', _pi), ('cable_body', _pi), ('cable_x', _pd), ('cable_ks', _pd), ('cable_kd', _pd), ('cable_L', _pd),
        ('chart', _pi),
    ]


class rmx_opts(C.Structure):
    _fields_ = [
        ('scheme', C.c_int32), ('nsteps', C.c_int32),
        ('h', C.c_double), ('tol', C.c_double), ('dxMax', C.c_double),
        ('iterMaxFactor', C.c_int32), ('iterLsMax', C.c_int32), ('linsolve', C.c_int32),
        ('ngpus', C.c_int32), ('tau_mode', C.c_int32), ('pcg_maxit', C.c_int32), ('pcg_tol', C.c_double),
    ]


class rmx_task_pointpos(C.Structure):
    _fields_ = [
        ('body', C.c_int32), ('reserved', C.c_int32),
        ('xlocal', C.c_double * 3),
        ('t_target', C.c_double), ('pscale', C.c_double), ('wreg', C.c_double), ('wpos', C.c_double),
    ]


# every symbol include/redmax_b200.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    'rmx_version', 'rmx_last_error', 'rmx_device_count', 'rmx_opts_default',
    'rmx_scene_create', 'rmx_scene_destroy', 'rmx_scene_nr', 'rmx_scene_nm',
    'rmx_rollout', 'rmx_rollout_resume', 'rmx_rollout_dev', 'rmx_rollout_adjoint', 'rmx_rollout_adjoint_dev',
    'rmx_adjoint_tape_bytes', 'rmx_eval', 'rmx_eval_newton', 'rmx_energies', 'rmx_linsolve_stats', 'rmx_debug_schedule', 'rmx_body_frames',
    'rmx_fp64_probe', 'rmx_host_register', 'rmx_host_unregister', 'rmx_eval_krylov', 'rmx_rollout_multi_dev',
]

_lib = None


class RmxError(RuntimeError):
    pass


def lib():
    """Load the CUDA library; fail loudly if it was not built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RmxError('%s is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                       '(redmax_b200 has no CPU fallback)' % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.rmx_version.restype = C.c_int
    L.rmx_last_error.restype = C.c_char_p
    L.rmx_device_count.restype = C.c_int
    L.rmx_opts_default.argtypes = [C.POINTER(rmx_opts), C.c_int32, C.c_int32]
    L.rmx_opts_default.restype = None
    L.rmx_scene_create.argtypes = [C.POINTER(rmx_scene_desc), C.POINTER(vp)]
    L.rmx_scene_destroy.argtypes = [vp]
    L.rmx_scene_destroy.restype = None
    L.rmx_scene_nr.argtypes = [vp]
    L.rmx_scene_nm.argtypes = [vp]
    L.rmx_rollout.argtypes = [vp, C.POINTER(rmx_opts), C.c_int64, vp, vp, vp, vp, vp, vp, vp]
    L.rmx_rollout_resume.argtypes = [vp, C.POINTER(rmx_opts), C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.rmx_rollout_dev.argtypes = [vp, C.POINTER(rmx_opts), C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.rmx_rollout_adjoint.argtypes = [vp, C.POINTER(rmx_opts), C.POINTER(rmx_task_pointpos), C.c_int64,
                                      vp, vp, vp, vp, vp, vp, vp, vp]
    L.rmx_rollout_adjoint_dev.argtypes = [vp, C.POINTER(rmx_opts), C.POINTER(rmx_task_pointpos), C.c_int64,
                                          vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.rmx_adjoint_tape_bytes.argtypes = [vp, C.POINTER(rmx_opts), C.c_int64]
    L.rmx_adjoint_tape_bytes.restype = C.c_int64
    L.rmx_eval.argtypes = [vp, vp, vp, vp, vp, C.c_double, C.c_double, vp, vp, vp, vp, vp]
    L.rmx_eval_newton.argtypes = [vp, vp, vp, vp, vp, C.c_double, C.c_double, vp, vp]
    L.rmx_rollout_multi_dev.argtypes = [vp, C.POINTER(rmx_opts), C.c_int32, vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp, C.c_int32]
    L.rmx_eval_krylov.argtypes = [vp, vp, vp, vp, vp, C.c_double, C.c_double, vp, vp, vp]
    L.rmx_energies.argtypes = [vp, C.c_int64, vp, vp, vp, vp]
    L.rmx_body_frames.argtypes = [vp, C.c_int64, vp, vp]
    L.rmx_debug_schedule.argtypes = [C.c_int64, C.c_int32, C.c_int64, C.c_int32, vp, vp]
    L.rmx_linsolve_stats.argtypes = [vp, C.POINTER(C.c_int64)]
    L.rmx_fp64_probe.argtypes = [_pd, _pd]
    L.rmx_host_register.argtypes = [vp, C.c_size_t]
    L.rmx_host_unregister.argtypes = [vp]
    _lib = L
    return L


def check(rc, what):
    if rc != RMX_OK:
        raise RmxError('%s failed (%d): %s' % (what, rc, lib().rmx_last_error().decode()))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr(a):
    """void* of a numpy array, a torch tensor (host or device), an int address, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, 'data_ptr'):
        return C.c_void_p(a.data_ptr())
    raise TypeError('cannot take the address of %r' % type(a))
