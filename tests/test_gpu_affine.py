"""GPU parity tests of the three affine drivers through the C ABI, in both
device representations of the constraint operator (sparse SDDMM / row-list path and dense DMMA-GEMM path)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FORCE = {"auto": 0, "dense": 1 | 4, "sparseA_denseS": 1 | 8, "sparse": 2 | 8}


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(1e-300, np.linalg.norm(b))


def _bqp(q):
    from oracle import generators as g
    d = np.load(os.path.join(GOLDEN, f"bqp_{q}_1.npz"))
    At, b, c, K = g.bqpmom(q, d["Q"], d["e"])
    return At, np.asarray(b.todense()).ravel() if hasattr(b, "todense") else np.asarray(b).ravel(), \
        (c / np.abs(c).max()), int(K["s"])


def _theta():
    from oracle import generators as g
    At, b, c, K = g.generate_hamming(7, [5, 6])
    return At, np.asarray(b.todense()).ravel() if hasattr(b, "todense") else np.asarray(b).ravel(), c, int(K["s"])


def _qs():
    from oracle import generators as g
    d = np.load(os.path.join(GOLDEN, "qs_c_10_1.npz"))
    At, b, c, K = g.qsmom(10, d["coe"])
    return At, np.asarray(b.todense()).ravel() if hasattr(b, "todense") else np.asarray(b).ravel(), c, int(K["s"])


def _dense_c(c):
    import scipy.sparse as sp
    return np.asarray(c.todense()).ravel() if sp.issparse(c) else np.asarray(c, dtype=np.float64).ravel()


def _point(kind, n, p, seed):
    rng = np.random.default_rng(seed)
    Y = rng.standard_normal((n, p))
    if kind == "unitdiag":
        Y /= np.linalg.norm(Y, axis=1, keepdims=True)
    elif kind == "unittrace":
        Y /= np.linalg.norm(Y)
    return Y, rng


PROBLEMS = {"bqp10": _bqp, "theta": _theta, "qs10": _qs}


@pytest.mark.parametrize("kind", ["unitdiag", "unittrace", "general"])
@pytest.mark.parametrize("prob,modes", [("bqp10", ["dense", "sparse", "sparseA_denseS"]), ("theta", ["auto", "dense"]),
                                        ("qs10", ["auto"])])
@pytest.mark.parametrize("p", [1, 5, 18])
def test_affine_closures_match_oracle(kind, prob, modes, p):
    """cost / grad / hess of ManiSDP_unitdiag.m:152-171, ManiSDP_unittrace.m:156-177, ManiSDP.m:149-165 at a random
    point with a non-trivial dual vector and penalty."""
    from manisdp_matlab_b200 import Handle
    from oracle.manisdp_ref import AffineProblem
    At, b, c, n = (_bqp(10) if prob == "bqp10" else PROBLEMS[prob]())
    m = At.shape[1]
    Y, rng = _point(kind, n, p, 10 * p + len(kind))
    U = rng.standard_normal((n, p))
    y = 0.3 * rng.standard_normal(m)
    sigma = 0.7
    ora = AffineProblem(kind, At.tocsc(), b, _dense_c(c), n, p, y, sigma)
    f0 = ora.cost(Y)
    g0 = ora.grad(Y)
    Ut = ora.M.proj(Y, U)
    H0 = ora.hess(Y, Ut)
    for mode in modes:
        with Handle(kind, n, At=At, b=b, c=c, force_mode=FORCE[mode]) as h:
            st = h.stats()
            if mode == "dense":
                assert (st.s_mode, st.a_mode) == (2, 2)
            if mode == "sparse":
                assert (st.s_mode, st.a_mode) == (1, 1)
            h.set_dual(y, sigma)
            h.set_Y(Y)
            f = h.cost()
            assert abs(f - f0) <= 1e-11 * max(1.0, abs(f0)), (mode, f, f0)
            g, gn = h.grad()
            assert _rel(g, g0) < 1e-11, mode
            assert abs(gn - np.linalg.norm(g0)) <= 1e-11 * np.linalg.norm(g0)
            assert _rel(h.project(U), Ut) < 1e-12
            assert _rel(h.hess(Ut), H0) < 1e-11, mode


def test_index_split_is_exact_on_large_linear_indices():
    """At row index r = j*n + i is split in 64-bit integers: a constraint touching X(n-1, n-1) of an n = 3000 matrix
    (r = 8 999 999 > 2^23, not exactly representable as an FP32 product path) must land on the right entry."""
    import scipy.sparse as sp
    from manisdp_matlab_b200 import Handle
    n, p = 3000, 3
    rows = np.array([(n - 1) * n + (n - 1), 0 * n + 1, 1 * n + 0, 2999 * n + 7, 7 * n + 2999], dtype=np.int64)
    cols = np.array([0, 1, 1, 2, 2])
    vals = np.array([2.0, 0.5, 0.5, 1.5, 1.5])
    At = sp.csc_matrix((vals, (rows, cols)), shape=(n * n, 3))
    b = np.array([1.0, 0.2, -0.4])
    c = sp.csc_matrix((np.array([1.0, 1.0]), (np.array([5 * n + 5, 9 * n + 9]), np.zeros(2, dtype=np.int64))),
                      shape=(n * n, 1))
    rng = np.random.default_rng(0)
    Y = rng.standard_normal((n, p))
    sigma = 2.0
    y = np.array([0.1, -0.2, 0.3])
    w = np.array([2.0 * Y[n - 1] @ Y[n - 1], Y[0] @ Y[1], 3.0 * Y[7] @ Y[2999]])
    r = w - b - y / sigma
    f0 = Y[5] @ Y[5] + Y[9] @ Y[9] + 0.5 * sigma * r @ r
    with Handle("general", n, At=At, b=b, c=c) as h:
        assert h.stats().a_mode == 1
        h.set_dual(y, sigma)
        h.set_Y(Y)
        assert abs(h.cost() - f0) <= 1e-12 * abs(f0)


@pytest.mark.parametrize("prob,mode", [("bqp10", "dense"), ("bqp10", "sparse"), ("theta", "auto"), ("qs10", "auto")])
def test_index_split_read_back_is_bit_exact(prob, mode):
    """BASELINE north_star: "A(YY') index handling must be bit-exact".  The (i, j) pairs the device kernels index with
    are read back through the C ABI and compared as INTEGER arrays with r mod n / r div n of At's CSC row indices
    (ManiSDP_unitdiag.m:153-155 applies A to vec(X) column-major: r = j*n + i)."""
    from manisdp_matlab_b200 import Handle
    At, b, c, n = (_bqp(10) if prob == "bqp10" else PROBLEMS[prob]())
    Ac = At.tocsc()
    Ac.sort_indices()
    r = Ac.indices.astype(np.int64)
    with Handle("unitdiag" if prob == "bqp10" else "general", n, At=Ac, b=b, c=c, force_mode=FORCE[mode]) as h:
        i, j = h.index_split()
    assert i.dtype == np.int64 and i.shape == r.shape
    assert np.array_equal(i, r % n) and np.array_equal(j, r // n)


def test_index_split_read_back_large_n():
    """same read-back where r exceeds 2^32: n = 100 000 (sparse representation), r up to n*n - 1 = 1e10 - 1"""
    import scipy.sparse as sp
    from manisdp_matlab_b200 import Handle
    n = 100000
    rng = np.random.default_rng(5)
    ii = rng.integers(0, n, 500)
    jj = rng.integers(0, n, 500)
    ii[0], jj[0] = n - 1, n - 1
    rows = np.concatenate([jj * n + ii, ii * n + jj]).astype(np.int64)
    cols = np.concatenate([np.arange(500), np.arange(500)])
    At = sp.csc_matrix((np.ones(1000), (rows, cols)), shape=(n * n, 500))
    At.sum_duplicates()
    At.sort_indices()
    c = sp.csc_matrix((np.array([1.0]), (np.array([0]), np.array([0]))), shape=(n * n, 1))
    with Handle("general", n, At=At, b=np.zeros(500), c=c) as h:
        i, j = h.index_split()
    r = At.indices.astype(np.int64)
    assert r.max() == n * n - 1 > 2 ** 32
    assert np.array_equal(i, r % n) and np.array_equal(j, r // n)


@pytest.mark.parametrize("kind,prob", [("unitdiag", "bqp10"), ("general", "qs10"), ("unittrace", "theta")])
@pytest.mark.parametrize("use_graph", [0, 1])
def test_affine_tr_iterates_match_oracle(kind, prob, use_graph):
    from manisdp_matlab_b200 import Handle
    from oracle.manisdp_ref import AffineProblem
    from oracle.manopt_rtr import trustregions
    At, b, c, n = PROBLEMS[prob]() if prob != "bqp10" else _bqp(10)
    m = At.shape[1]
    p = 4
    Y0, rng = _point(kind, n, p, 99)
    y = 0.1 * rng.standard_normal(m)
    sigma = 5.0
    ora = AffineProblem(kind, At.tocsc(), b, _dense_c(c), n, p, y, sigma)
    res = trustregions(ora, Y0.copy(), maxiter=6, maxinner=15, tolgradnorm=1e-10)
    with Handle(kind, n, At=At, b=b, c=c) as h:
        h.set_dual(y, sigma)
        h.set_Y(Y0)
        info = h.tr_solve(maxiter=6, maxinner=15, tolgradnorm=1e-10, use_graph=use_graph)
        log = h.tr_log()
        Yd = h.get_Y()
    assert [r.numinner for r in log] == [r.numinner for r in res.info]
    assert [r.accepted for r in log] == [int(r.accepted) for r in res.info]
    assert [r.stop_inner for r in log] == [r.stop_inner for r in res.info]
    assert abs(info.cost - res.cost) <= 1e-9 * max(1.0, abs(res.cost))
    assert abs(info.gradnorm - res.info[-1].gradnorm) <= 1e-6 * max(res.info[-1].gradnorm, 1e-8)
    assert _rel(Yd, res.x) < 1e-7


def test_kkt_matches_dense_formulas():
    """residues, dual update and dual slack of ManiSDP_unitdiag.m:59-71 against a dense NumPy evaluation"""
    from manisdp_matlab_b200 import Handle
    for kind, prob in [("unitdiag", "bqp10"), ("unittrace", "theta"), ("general", "qs10")]:
        At, b, c, n = PROBLEMS[prob]() if prob != "bqp10" else _bqp(10)
        cd = _dense_c(c)
        m = At.shape[1]
        Y, rng = _point(kind, n, 6, 7)
        y = 0.05 * rng.standard_normal(m)
        sigma = 3.0
        X = Y @ Y.T
        x = X.reshape(-1, order="F")
        Axb = At.T @ x - b
        y1 = y - sigma * Axb
        eS = (cd - At @ y1).reshape(n, n, order="F")
        if kind == "unitdiag":
            z = np.sum(X * eS, axis=0)
            S = eS - np.diag(z)
            by = b @ y1 + z.sum()
        elif kind == "unittrace":
            z = np.sum(X * eS)
            S = eS - z * np.eye(n)
            by = b @ y1 + z
        else:
            S = eS
            by = b @ y1
        dS = np.linalg.eigvalsh(S)
        obj = cd @ x
        with Handle(kind, n, At=At, b=b, c=c) as h:
            h.set_dual(y, sigma)
            h.set_Y(Y)
            k = h.kkt(8, 1e-11, 1)
            ynew, _ = h.get_dual()
            vals, vecs = h.get_eigs(min(8, n))
        assert abs(k.obj - obj) <= 1e-11 * max(1, abs(obj))
        assert abs(k.pinf - np.linalg.norm(Axb) / (1 + np.linalg.norm(b))) <= 1e-11
        assert _rel(ynew, y1) < 1e-12
        assert abs(k.by - by) <= 1e-10 * max(1, abs(by))
        assert abs(k.gap - abs(obj - by) / (abs(by) + abs(obj) + 1)) <= 1e-10
        assert np.allclose(vals, dS[:len(vals)], atol=1e-7 * (1 + abs(dS).max()))
        assert abs(k.dinf - max(0, -dS[0]) / (1 + dS[-1])) <= 1e-3 * abs(k.dinf) + 1e-9


@pytest.mark.parametrize("q", [10, 20])
def test_bqp_full_solve(q):
    """BASELINE config 2 family: optimum (rel 1e-6) of the oracle / brute force, all KKT residues <= 1e-8."""
    from manisdp_matlab_b200 import ManiSDP_unitdiag
    gold = json.load(open(os.path.join(GOLDEN, "oracle_outputs.json")))
    At, b, c, n = _bqp(q)
    X, obj, data = ManiSDP_unitdiag(At, b, c, {"s": n}, dict(verbose=False))
    assert data["status"] == 0
    assert max(data["gap"], data["pinf"], data["dinf"]) < 1e-8
    assert abs(obj - gold[f"bqp_{q}_1_opt"]["obj_scaled"]) <= 1e-6 * abs(gold[f"bqp_{q}_1_opt"]["obj_scaled"])
    if q == 10:
        from oracle import generators as g
        d = np.load(os.path.join(GOLDEN, "bqp_10_1.npz"))
        _, _, c0, _ = g.bqpmom(10, d["Q"], d["e"])
        assert abs(obj * np.abs(c0).max() - gold["bqp_10_1_bruteforce"]) <= 1e-6 * abs(gold["bqp_10_1_bruteforce"])


def test_qsphere_full_solve():
    """BASELINE config 3 family (qs_c_10_1 through ManiSDP, options of example_qsphere.m:18-27)."""
    from manisdp_matlab_b200 import ManiSDP
    gold = json.load(open(os.path.join(GOLDEN, "oracle_outputs.json")))["qs_c_10_1_opt"]
    At, b, c, n = _qs()
    X, obj, data = ManiSDP(At, b, c, {"s": n}, dict(verbose=False, tol=1e-8, theta=1e-2, tau1=0.02))
    assert max(data["gap"], data["pinf"], data["dinf"]) < 1e-8
    assert abs(obj - gold["obj"]) <= 1e-6 * abs(gold["obj"])


def test_theta_full_solve():
    """BASELINE config 4 family: Hamming(7,[5,6]) through ManiSDP_unittrace with example_theta.m:48-55 options;
    theta = 42.6667 (DIMACS)."""
    from manisdp_matlab_b200 import ManiSDP_unittrace
    At, b, c, n = _theta()
    X, obj, data = ManiSDP_unittrace(At, b, c, {"s": n}, dict(verbose=False, tol=1e-6, sigma0=1e5, sigma_max=1e8,
                                                             line_search=1))
    assert max(data["gap"], data["pinf"], data["dinf"]) < 1e-6
    assert abs(obj - (-128.0 / 3.0)) <= 1e-4 * 42.67


@pytest.mark.parametrize("kind", ["unitdiag", "unittrace", "general"])
def test_affine_closures_at_widths_beyond_512(kind):
    """factor widths 512 < p <= 1024 (row groups with 16 vectors per lane): cost / grad / hess and the manifold operations
    against the oracle on BQP-20 (n = 211) at p = 600 -- the reference has no width limit; BQP d >= 120 needs p > 512"""
    from manisdp_matlab_b200 import Handle
    from oracle.manisdp_ref import AffineProblem
    At, b, c, n = _bqp(20)
    m, p = At.shape[1], 600
    Y, rng = _point(kind, n, p, 5)
    U = rng.standard_normal((n, p))
    y = 0.3 * rng.standard_normal(m)
    sigma = 0.7
    ora = AffineProblem(kind, At.tocsc(), b, _dense_c(c), n, p, y, sigma)
    f0 = ora.cost(Y)
    g0 = ora.grad(Y)
    Ut = ora.M.proj(Y, U)
    H0 = ora.hess(Y, Ut)
    R0 = ora.M.retr(Y, 0.1 * Ut)
    for mode in ("auto", "sparse"):
        with Handle(kind, n, At=At, b=b, c=c, force_mode=FORCE[mode]) as h:
            h.set_dual(y, sigma)
            h.set_Y(Y)
            f = h.cost()
            G, gn = h.grad()
            P = h.project(U)
            Hd = h.hess(Ut)
            R = h.retract(0.1 * Ut)
        assert abs(f - f0) <= 1e-11 * max(1.0, abs(f0))
        assert _rel(G, g0) < 1e-11 and _rel(P, Ut) < 1e-12 and _rel(Hd, H0) < 1e-10 and _rel(R, R0) < 1e-12
    with Handle(kind, n, At=At, b=b, c=c) as h:
        with pytest.raises(Exception, match="1024"):
            h.set_Y(np.zeros((n, 1030)))
