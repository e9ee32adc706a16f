"""ctypes binding of oracle/cspn_oracle.c.  TEST INFRASTRUCTURE ONLY (see the C file's header)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'libcspn_oracle.so')
_lib = None


def build(force=False):
    src = os.path.join(HERE, 'cspn_oracle.c')
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', HERE, '-s', '-B' if force else '-s'])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        fp = ctypes.POINTER(ctypes.c_float)
        _lib.cspn2d_oracle_f32.argtypes = [fp, fp, fp, fp] + [ctypes.c_int] * 8
        _lib.cspn2d_oracle_f32.restype = ctypes.c_int
        _lib.cspn3d_oracle_f32.argtypes = [fp, fp, fp] + [ctypes.c_int] * 8
        _lib.cspn3d_oracle_f32.restype = ctypes.c_int
        _lib.cspn_oracle_max_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if a is not None else None


def max_threads():
    return int(lib().cspn_oracle_max_threads())


def cspn2d(guidance, blur_depth, sparse_depth=None, prop_time=24, norm_type='8sum', nthreads=0):
    g = np.ascontiguousarray(guidance, dtype=np.float32)
    d = np.ascontiguousarray(blur_depth, dtype=np.float32)
    s = None if sparse_depth is None else np.ascontiguousarray(sparse_depth, dtype=np.float32)
    B, gch, H, W = g.shape
    C = d.shape[1]
    out = np.empty_like(d)
    rc = lib().cspn2d_oracle_f32(_p(g), _p(d), _p(s), _p(out), B, C, H, W, gch, int(prop_time),
                                 int('abs' in norm_type), int(nthreads))
    if rc != 0:
        raise MemoryError('oracle allocation failed')
    return out


_MODE3 = {'26sum': 0, '26sum_abs': 1, 'paddle': 2}


def cspn3d(guidance, feat, prop_time=12, norm_type='26sum_abs', nthreads=0):
    g = np.ascontiguousarray(guidance, dtype=np.float32)
    f = np.ascontiguousarray(feat, dtype=np.float32)
    B, C, D, H, W = f.shape
    assert g.shape == (B, 26, D, H, W)
    out = np.empty_like(f)
    rc = lib().cspn3d_oracle_f32(_p(g), _p(f), _p(out), B, C, D, H, W, int(prop_time), _MODE3[norm_type],
                                 int(nthreads))
    if rc != 0:
        raise MemoryError('oracle allocation failed')
    return out
