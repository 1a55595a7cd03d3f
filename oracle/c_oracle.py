"""ctypes access to oracle/_build/libfemoracle.so (C + OpenMP CG on CSR) -- test / baseline infrastructure only."""
import ctypes as C
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libfemoracle.so")
_lib = C.CDLL(_PATH)   # raises OSError when not built: callers fall back to the numpy oracle
_lib.oracle_threads.restype = C.c_int
_lib.oracle_cg_csr.restype = C.c_int
_P = C.c_void_p


def threads():
    return int(_lib.oracle_threads())


def _ptr(a):
    return _P(a.ctypes.data)


def csr_matvec(crow, col, val, x):
    crow, col = np.ascontiguousarray(crow, np.int64), np.ascontiguousarray(col, np.int64)
    val, x = np.ascontiguousarray(val, np.float64), np.ascontiguousarray(x, np.float64).reshape(-1)
    y = np.empty(crow.size - 1)
    _lib.oracle_csr_matvec(C.c_int64(y.size), _ptr(crow), _ptr(col), _ptr(val), _ptr(x), _ptr(y))
    return y


def cg_csr(crow, col, val, F, fixed_dofs, u_init=None, tol=1e-10, max_iter=1000, eps=1e-30):
    """The reference's projected CG loop on a CSR operator; returns (u, iterations, status)."""
    crow, col = np.ascontiguousarray(crow, np.int64), np.ascontiguousarray(col, np.int64)
    val = np.ascontiguousarray(val, np.float64)
    Ff = np.ascontiguousarray(F, np.float64).reshape(-1)
    n = Ff.size
    free = np.ones(n, np.uint8)
    free[np.asarray(fixed_dofs, np.int64)] = 0
    u = np.zeros(n) if u_init is None else np.ascontiguousarray(u_init, np.float64).reshape(-1).copy()
    work = np.empty(3 * n)
    st = C.c_int(0)
    it = _lib.oracle_cg_csr(C.c_int64(n), _ptr(crow), _ptr(col), _ptr(val), _ptr(Ff), _ptr(free), _ptr(u), _ptr(work), C.c_double(tol),
                            C.c_int(max_iter), C.c_double(eps), C.byref(st))
    return u.reshape(np.asarray(F).shape), int(it), {0: "converged", 1: "breakdown", 2: "maxiter"}[st.value]
