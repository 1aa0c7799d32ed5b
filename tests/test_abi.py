"""CPU: the C-ABI library loads and exports every symbol include/manisdp_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "manisdp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(manisdp_[a-zA-Z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(engine_lib):
    names = _declared()
    assert len(names) >= 28
    for n in names:
        assert hasattr(engine_lib, n), f"{n} declared in include/manisdp_b200.h but not exported"


def test_binding_covers_header(engine_lib):
    from manisdp_matlab_b200 import _lib
    assert set(_declared()) == set(_lib.SIGNATURES)


def test_struct_sizes_match_header():
    # sizes the MEX gateway / any FFI must agree on (LP64)
    from manisdp_matlab_b200 import _lib
    assert ctypes.sizeof(_lib.Problem) == 8 + 16 + 3 * 8 + 3 * 8 + 8 + 8 + 8 + 8 + 8 + 16 + 8 + 8 + 8 + 8 + 6 * 8
    assert ctypes.sizeof(_lib.TrOptions) == 16 + 7 * 8
    assert ctypes.sizeof(_lib.TrInfo) == 4 * 8 + 8 + 16
    assert ctypes.sizeof(_lib.TrIter) == 5 * 8 + 16
    assert ctypes.sizeof(_lib.KktInfo) == 8 * 8 + 8 + 8 + 8


def test_version_and_error_paths(engine_lib):
    assert engine_lib.manisdp_version() == 100
    # argument errors are reported without touching a GPU
    assert engine_lib.manisdp_create(None, None) == -1
    assert b"null" in engine_lib.manisdp_last_error(None)


def test_no_gpu_means_loud_failure(engine_lib):
    """On a CPU-only box create() must fail with MANISDP_E_CUDA -- never fall back."""
    import torch
    if torch.cuda.is_available():
        return
    import scipy.sparse as sp
    from manisdp_matlab_b200 import _lib
    try:
        _lib.Handle("onlyunitdiag", 4, C_csc=sp.identity(4, format="csc"))
    except _lib.EngineError as e:
        assert "(-2)" in str(e)
    else:
        raise AssertionError("create succeeded without a GPU")


def test_host_small_eigensolver(engine_lib):
    from manisdp_matlab_b200 import _lib
    rng = np.random.default_rng(0)
    for n in [1, 2, 7, 36, 48, 64, 65, 130, -40, -90]:  # negative: a degenerate spectrum (three distinct eigenvalues)
        if n > 0:
            A = rng.standard_normal((n, n))
            A = A + A.T
        else:
            n = -n
            Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
            A = (Q * rng.choice([-1.0, 0.0, 2.5], size=n)) @ Q.T
            A = 0.5 * (A + A.T)
        w = np.empty(n)
        V = np.empty((n, n))
        assert engine_lib.manisdp_test_sym_eig(_lib._pf(np.ascontiguousarray(A)), n, _lib._pf(w), _lib._pf(V)) == 0
        assert np.allclose(w, np.linalg.eigvalsh(A), atol=1e-11 * max(1, n))
        assert np.abs(A @ V - V * w).max() < 1e-11 * max(1, n)
        assert np.abs(V.T @ V - np.eye(n)).max() < 1e-12 * max(1, n)
        w2 = np.empty(n)  # eigenvalues-only path (rank step): V = NULL
        assert engine_lib.manisdp_test_sym_eig(_lib._pf(np.ascontiguousarray(A)), n, _lib._pf(w2), None) == 0
        assert np.allclose(w2, np.linalg.eigvalsh(A), atol=1e-11 * max(1, n))
