"""The C-ABI shared library loads on a CPU-only box and exports exactly what include/v3s.h declares
(no compute calls here: those need a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "v3s.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(v3s_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from v3s import _lib
    names = declared_symbols()
    assert len(names) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    unbound = [n for n in names if n not in _lib.SIGNATURES]
    assert not unbound, unbound
    undeclared = [n for n in _lib.SIGNATURES if n not in names]
    assert not undeclared, undeclared
    assert lib.v3s_version() == 100


def test_error_plumbing_without_gpu():
    from v3s import _lib
    # argument validation happens before any CUDA call: a bad shape returns non-zero with a message
    rc = _lib.lib.v3s_block_matvec(None, 3, 1, 10, None, 99, None, None, None)
    assert rc != 0 and "block size" in _lib.last_error()
    rc = _lib.lib.v3s_sparse_block_matvec(None, 0, 1, 10, None, 8, None, None)
    assert rc != 0 and "operator descriptor" in _lib.last_error()
    assert _lib.lib.v3s_xf_floats(130) == 2048
    assert _lib.lib.v3s_synth_workspace_bytes(64) == 64 * 64 * 16


def test_sym_eig_host_against_numpy():
    """v3s_sym_eig_host (Householder + QL, the Ritz extraction of v3s_eig_topk and the N <= 600 eigen path): a host
    routine, so it is checked here on the CPU box."""
    import ctypes as C

    import numpy as np
    from v3s import _lib
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 8, 33, 128, 200):
        A = rng.standard_normal((n, n))
        A = (A + A.T) / 2
        if n == 33:                                   # a degenerate cluster (lambda = 1 eleven times) plus a spread
            A = np.diag(np.r_[np.ones(11), rng.random(22)]) + 1e-9 * A
        a, w = np.ascontiguousarray(A.copy()), np.zeros(n)
        rc = _lib.lib.v3s_sym_eig_host(a.ctypes.data_as(C.POINTER(C.c_double)), n, w.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        assert np.max(np.abs(w - np.linalg.eigvalsh(A))) < 1e-12 * max(1.0, np.abs(A).max())
        assert np.max(np.abs(A @ a - a * w[None, :])) < 1e-12 * max(1.0, np.abs(A).max())
        assert np.max(np.abs(a.T @ a - np.eye(n))) < 1e-12
    # the operator of a disconnected kNN graph (the reference's default run has 182 components): a permuted block-diagonal
    # normalised Markov matrix -- eigenvalue 1 once per component, rank-deficient blocks with zero clusters
    def markov_blocks(nb, bs):
        n = nb * bs
        M = np.zeros((n, n))
        for i in range(nb):
            W = rng.random((bs, bs))
            W = W + W.T
            if i % 3 == 0:
                W[:] = 1.0                              # rank one: eigenvalue 0 with multiplicity bs - 1
            d = W.sum(1)
            M[i * bs:(i + 1) * bs, i * bs:(i + 1) * bs] = W / np.sqrt(np.outer(d, d))
        p = rng.permutation(n)
        return M[np.ix_(p, p)]
    for A in (markov_blocks(64, 8), markov_blocks(182, 3), np.zeros((40, 40)), np.ones((64, 64)), np.eye(50)):
        n = A.shape[0]
        a, w = np.ascontiguousarray(A.copy()), np.zeros(n)
        rc = _lib.lib.v3s_sym_eig_host(a.ctypes.data_as(C.POINTER(C.c_double)), n, w.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0, _lib.last_error()
        assert np.max(np.abs(w - np.linalg.eigvalsh(A))) < 1e-12 * max(1.0, np.abs(A).max())
        assert np.max(np.abs(A @ a - a * w[None, :])) < 1e-12 * max(1.0, np.abs(A).max())
        assert np.max(np.abs(a.T @ a - np.eye(n))) < 1e-12
    assert _lib.lib.v3s_sym_eig_host(None, 3, None) != 0 and "bad arguments" in _lib.last_error()
