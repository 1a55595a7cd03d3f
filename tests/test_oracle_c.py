"""The compiled (C + pthreads) CG used as the multi-threaded CPU baseline must agree with the numpy oracle."""
import numpy as np
import pytest

from conftest import PKG  # noqa: F401  (path setup)


def test_c_oracle_matches_numpy_oracle():
    c_oracle = pytest.importorskip("oracle.c_oracle", reason="oracle/_build not built (run __graft_entry__.build())")
    from femb200 import meshgen
    from oracle import fem_oracle as O
    c, t = meshgen.kuhn_cube(8, jitter=0.1)
    c, t = c.numpy(), t.numpy()
    Ke = O.c3d4_poisson_K(c, t)
    crow, col, val, n = O.assemble_csr(Ke, t, 1, c.shape[0])
    load = np.bincount(t.reshape(-1), weights=np.repeat(O.tet_volumes(c, t) / 4, 4), minlength=c.shape[0]).reshape(-1, 1)
    fixed = np.flatnonzero(c[:, 2] == 0)
    u1, it1, s1 = O.stable_cg(Ke, t, load, fixed, tol=1e-9, ndof=1)
    u2, it2, s2 = c_oracle.cg_csr(crow, col, val, load, fixed, tol=1e-9)
    assert s1 == s2 == "converged" and abs(it1 - it2) <= 1
    assert np.abs(u1 - u2).max() <= 1e-8 * np.abs(u1).max()
    x = np.random.default_rng(0).standard_normal(n)
    assert np.abs(c_oracle.csr_matvec(crow, col, val, x) - O.csr_matvec(crow, col, val, x)).max() < 1e-12
    u3, it3, s3 = c_oracle.cg_csr(crow, col, val, load, fixed, tol=1e-9, max_iter=7)
    assert s3 == "maxiter" and it3 == 7
