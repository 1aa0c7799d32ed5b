"""Known answers from the SDPLIB table shipped in the reference tree (data/sdplib/README:21-112), through the SDPA
reader of manisdp_matlab_b200.problems.  SeDuMi sign convention: obj = -(SDPLIB optimal value).
CPU part pins the oracle; GPU part checks the engine against the same table."""
import os

import numpy as np
import pytest

SDPLIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdplib")
TABLE = {"mcp100": 226.1574, "mcp124-1": 141.9905, "theta1": 23.0}


def _load(name):
    from manisdp_matlab_b200 import problems as P
    return P.read_sdpa(os.path.join(SDPLIB, name + ".dat-s"))


def test_reader_shapes_and_symmetry():
    At, b, c, K = _load("mcp100")
    n = K["s"]
    assert (n, At.shape[1], len(b)) == (100, 100, 100)
    Cm = c.toarray().reshape(n, n, order="F")
    assert np.array_equal(Cm, Cm.T)
    # constraint k is X_kk = 1
    A0 = At[:, 0].toarray().reshape(n, n, order="F")
    assert A0[0, 0] == 1.0 and np.count_nonzero(A0) == 1 and np.all(b == 1.0)
    At, b, c, K = _load("theta1")
    assert (K["s"], At.shape[1]) == (50, 104)
    A1 = At[:, 1].toarray().reshape(50, 50, order="F")
    assert np.array_equal(A1, A1.T) and np.count_nonzero(A1) == 2


def test_oracle_mcp100_matches_sdplib():
    from oracle import manisdp_ref as ref
    At, b, c, K = _load("mcp100")
    n = K["s"]
    C = c.toarray().reshape(n, n, order="F")
    X, obj, data = ref.ManiSDP_onlyunitdiag(C, dict(p0=10, seed=0))
    assert data["dinf"] < 1e-8
    assert abs(-obj - TABLE["mcp100"]) < 2e-4


def test_oracle_theta1_matches_sdplib():
    from oracle import manisdp_ref as ref
    At, b, c, K = _load("theta1")
    X, obj, data = ref.ManiSDP(At, b, c, K, dict(seed=0, sigma0=1e1, sigma_min=1e0, TR_maxiter=10, TR_maxinner=50))
    assert max(data["gap"], data["pinf"], data["dinf"]) < 1e-8
    assert abs(-obj - TABLE["theta1"]) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mcp100", "mcp124-1"])
def test_gpu_mcp_through_both_drivers(name):
    import scipy.sparse as sp
    from manisdp_matlab_b200 import ManiSDP, ManiSDP_onlyunitdiag
    At, b, c, K = _load(name)
    n = K["s"]
    C = sp.csc_matrix(c.toarray().reshape(n, n, order="F"))
    X, obj, data = ManiSDP_onlyunitdiag(C, dict(p0=10, verbose=False))
    assert data["dinf"] < 1e-8 and abs(-obj - TABLE[name]) < 2e-4
    # the same problem as a general affine SDP (diagonal constraints in At), Euclidean manifold + AL
    X2, obj2, d2 = ManiSDP(At, b, c, K, dict(verbose=False, sigma0=1e1, sigma_min=1e0, TR_maxiter=10, TR_maxinner=50))
    assert max(d2["gap"], d2["pinf"], d2["dinf"]) < 1e-8
    assert abs(-obj2 - TABLE[name]) < 2e-4


@pytest.mark.gpu
def test_gpu_theta1():
    from manisdp_matlab_b200 import ManiSDP
    At, b, c, K = _load("theta1")
    X, obj, data = ManiSDP(At, b, c, K, dict(verbose=False, sigma0=1e1, sigma_min=1e0, TR_maxiter=10, TR_maxinner=50))
    assert max(data["gap"], data["pinf"], data["dinf"]) < 1e-8
    assert abs(-obj - TABLE["theta1"]) < 1e-5
