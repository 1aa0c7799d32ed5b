"""GPU parity tests of the dual approach (src/dual/ManiDSDP_unitdiag.m:28-194, Riemannian ADMM on the
SOS form) through the C ABI (kind MANISDP_DUAL_UNITDIAG), against oracle/manisdp_ref.py::DualProblem / ManiDSDP_unitdiag."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FORCE = {"auto": 0, "sparseA": 8}


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(1e-300, np.linalg.norm(b))


def _sos(q):
    """example/dual/example_bqp_dual.m:19-33: SOS data of BQP q with the free variable of the constant term"""
    import scipy.sparse as sp
    from instances import generators as G
    d = np.load(os.path.join(GOLDEN, f"bqp_{q}_1.npz"))
    A, b, dAAt, mb = G.bqpsos(d["Q"], d["e"], q)
    v = np.zeros((A.shape[0], 1))
    v[0] = 1.0
    A2 = sp.hstack([sp.csr_matrix(v), A]).tocsr()
    c = np.concatenate([[1.0], np.zeros(mb * mb)])
    maxb = float(np.abs(b).max())
    return A2, b / maxb, c, {"f": 1, "s": mb}, dAAt, maxb, d


def _handle(A2, b, c, K, dAAt, **kw):
    from manisdp_matlab_b200 import Handle
    nf = K["f"]
    A2 = A2.tocsc()
    return Handle("dual_unitdiag", K["s"], At=A2[:, nf:].T.tocsc(), b=b, c=c[nf:], dAAt=dAAt, B=A2[:, :nf], cf=c[:nf], **kw)


def _state(n, m, nf, rng):
    """a multiplier x in the range of I - P is not required by the closures themselves, but the engine's cost uses
    P x = 0 (true for every x the ADMM iteration produces); tests build x accordingly"""
    M = rng.standard_normal((n, n))
    return 0.05 * (M + M.T).reshape(-1), 0.1 * rng.standard_normal(nf)


def _project_out(A, dAAt, x):
    """x - P x with P = A' D^-1 A"""
    return x - A.T @ ((A @ x) / dAAt)


@pytest.mark.parametrize("mode", ["auto", "sparseA"])
@pytest.mark.parametrize("p", [3, 14])
def test_dual_closures_match_oracle(mode, p):
    """cost / grad / hess of ManiDSDP_unitdiag.m:171-191 at a random point with non-trivial multipliers"""
    from oracle.manisdp_ref import DualProblem
    A2, b, c, K, dAAt, _, _ = _sos(10)
    n, nf, m = K["s"], K["f"], A2.shape[0]
    A, B = A2[:, nf:].tocsr(), A2[:, :nf].tocsr()
    rng = np.random.default_rng(40 + p)
    x, w = _state(n, m, nf, rng)
    x = _project_out(A, dAAt, x)
    sigma = 0.37
    Y = rng.standard_normal((n, p))
    Y /= np.linalg.norm(Y, axis=1, keepdims=True)
    ora = DualProblem(A, B, b, c[nf:], c[:nf], n, p, dAAt, x, w, sigma)
    f0 = ora.cost(Y)
    g0 = ora.grad(Y)
    U = ora.M.proj(Y, rng.standard_normal((n, p)))
    H0 = ora.hess(Y, U)
    with _handle(A2, b, c, K, dAAt, force_mode=FORCE[mode]) as h:
        assert h.stats().s_mode == 2 and h.stats().a_mode == (2 if mode == "auto" else 1)
        h.set_sigma(sigma)
        h.dual_set_state(x, w)
        h.set_Y(Y)
        f = h.cost()
        G, gn = h.grad()
        Hd = h.hess(U)
        xs, ws = h.dual_state()
    assert np.array_equal(xs, x) and np.array_equal(ws, w)
    assert abs(f - f0) <= 1e-11 * max(1.0, abs(f0)), (f, f0)
    assert _rel(G, g0) < 1e-11
    assert _rel(Hd, H0) < 1e-10
    assert abs(gn - np.linalg.norm(g0)) <= 1e-11 * max(1.0, gn)


@pytest.mark.parametrize("use_graph", [1, 0])
def test_dual_tr_iterates_match_oracle(use_graph):
    from oracle.manisdp_ref import DualProblem
    from oracle.manopt_rtr import trustregions
    A2, b, c, K, dAAt, _, _ = _sos(10)
    n, nf, m = K["s"], K["f"], A2.shape[0]
    A, B = A2[:, nf:].tocsr(), A2[:, :nf].tocsr()
    rng = np.random.default_rng(9)
    x, w = _state(n, m, nf, rng)
    x = _project_out(A, dAAt, x)
    sigma, p = 0.05, 6
    Y0 = rng.standard_normal((n, p))
    Y0 /= np.linalg.norm(Y0, axis=1, keepdims=True)
    ora = DualProblem(A, B, b, c[nf:], c[:nf], n, p, dAAt, x, w, sigma)
    res = trustregions(ora, Y0.copy(), maxiter=6, maxinner=15, tolgradnorm=1e-10)
    with _handle(A2, b, c, K, dAAt) as h:
        h.set_sigma(sigma)
        h.dual_set_state(x, w)
        h.set_Y(Y0)
        info = h.tr_solve(maxiter=6, maxinner=15, tolgradnorm=1e-10, use_graph=use_graph)
        log = h.tr_log()
        Yd = h.get_Y()
    assert [r.numinner for r in log] == [r.numinner for r in res.info]
    assert [r.accepted for r in log] == [int(r.accepted) for r in res.info]
    assert [r.stop_inner for r in log] == [r.stop_inner for r in res.info]
    assert abs(info.cost - res.cost) <= 1e-9 * max(1.0, abs(res.cost))
    assert _rel(Yd, res.x) < 1e-7


@pytest.mark.parametrize("q", [10, 20])
def test_dual_admm_step_matches_dense_formulas(q):
    """ManiDSDP_unitdiag.m:71-88 (y, As, Af, x, w, eX, z, obj, pinf, gap, dinf) against a dense NumPy evaluation;
    q = 20 (n = 211) goes through the LOBPCG eigen step, q = 10 (n = 56) through the small dense one"""
    A2, b, c, K, dAAt, _, _ = _sos(q)
    n, nf, m = K["s"], K["f"], A2.shape[0]
    A, B = A2[:, nf:].tocsr(), A2[:, :nf].tocsr()
    cp, cf = c[nf:], c[:nf]
    rng = np.random.default_rng(4)
    x, w = _state(n, m, nf, rng)
    x = _project_out(A, dAAt, x)
    sigma, p = 0.8, 5
    Y = rng.standard_normal((n, p))
    Y /= np.linalg.norm(Y, axis=1, keepdims=True)
    S = Y @ Y.T
    sc = S.reshape(-1, order="F") - cp
    y = (A @ sc) / dAAt
    As = A.T @ y - sc
    Af = B.T @ y - cf
    pinf = (np.linalg.norm(As) + np.linalg.norm(Af)) / (1 + np.linalg.norm(c))
    by = float(b @ y)
    x1 = x - sigma * As
    w1 = w - sigma * Af
    bA = A.T @ (b / dAAt)
    eX = (x1 + bA).reshape(n, n, order="F")
    z = np.sum(S * eX, axis=0)
    dX = np.linalg.eigvalsh(eX - np.diag(z))
    obj = float(cp @ eX.reshape(-1, order="F") + cf @ w1 + z.sum())
    with _handle(A2, b, c, K, dAAt) as h:
        h.set_sigma(sigma)
        h.dual_set_state(x, w)
        h.set_Y(Y)
        k = h.kkt(8, 1e-10, 1)
        yd, _ = h.get_dual()
        xd, wd = h.dual_state()
        vals, vecs = h.get_eigs(min(8, k.nneg if k.nneg > 0 else 1))
    assert _rel(yd, y) < 1e-12
    assert _rel(xd, x1) < 1e-12 and _rel(wd, w1) < 1e-12
    assert abs(k.pinf - pinf) <= 1e-12 * max(1.0, pinf)
    assert abs(k.by - by) <= 1e-12 * max(1.0, abs(by))
    assert abs(k.obj - obj) <= 1e-11 * max(1.0, abs(obj))
    assert abs(k.gap - abs(obj - by) / (1 + abs(obj) + abs(by))) <= 1e-12
    assert abs(k.lam_min - dX[0]) <= 1e-7 * max(1.0, abs(dX[0]))
    assert abs(k.dinf - max(0.0, -dX[0]) / (1 + abs(dX[-1]))) <= 1e-6
    assert k.nneg == min(8, int(np.sum(dX < 0)))


def test_dual_full_solve_bqp10_reaches_the_exhaustive_minimum():
    """oracle-free pin: the SOS bound of BQP-10 is tight, so ManiDSDP's optimum times max|b| equals the brute-force
    minimum (options of example/dual/example_bqp_dual.m:22-36)"""
    from instances import generators as G
    from manisdp_matlab_b200 import ManiDSDP_unitdiag
    A2, b, c, K, dAAt, maxb, d = _sos(10)
    X, obj, data = ManiDSDP_unitdiag(A2, b, c, K, dict(dAAt=dAAt, tol=1e-8, line_search=1, verbose=False))
    assert data["status"] == 0 and max(data["gap"], data["pinf"], data["dinf"]) < 1e-8
    assert abs(obj * maxb - G.bqp_bruteforce(d["Q"], d["e"])) <= 1e-6 * abs(obj * maxb)
    assert np.allclose(np.diag(data["S"]), 1.0, atol=1e-12)
    assert np.linalg.eigvalsh(X)[0] > -1e-6


@pytest.mark.parametrize("line_search", [0, 1])
def test_dual_full_solve_bqp20_matches_oracle_and_primal(line_search):
    """BQP-20 (n = 211, m = 6196): same optimum as the oracle's dual restatement and as the primal moment relaxation
    solved by ManiSDP_unitdiag on the same instance (strong duality)"""
    from instances import generators as G
    from manisdp_matlab_b200 import ManiDSDP_unitdiag, ManiSDP_unitdiag
    from oracle.manisdp_ref import ManiDSDP_unitdiag as ref_dual
    A2, b, c, K, dAAt, maxb, d = _sos(20)
    opts = dict(dAAt=dAAt, tol=1e-8, line_search=line_search, verbose=False)
    _, obj_ref, dref = ref_dual(A2, b, c, K, opts)
    assert dref["status"] == 0
    _, obj, data = ManiDSDP_unitdiag(A2, b, c, K, opts)
    assert data["status"] == 0 and max(data["gap"], data["pinf"], data["dinf"]) < 1e-8
    assert abs(obj - obj_ref) <= 1e-6 * max(1.0, abs(obj_ref))
    At, bp, cp, Kp = G.bqpmom(20, d["Q"], d["e"])
    mc = float(np.abs(cp).max())
    bp = np.asarray(bp.todense()).ravel() if hasattr(bp, "todense") else np.asarray(bp).ravel()
    _, obj_p, dp = ManiSDP_unitdiag(At, bp, cp / mc, Kp, dict(tol=1e-8, verbose=False))
    assert dp["status"] == 0
    assert abs(obj * maxb - obj_p * mc) <= 1e-5 * abs(obj_p * mc)


def test_dual_bqp60_reaches_the_baseline_optimum():
    """config 2's instance (data/bqp_{Q,e}_60_1.txt, n = 1831) through the DUAL driver at its stated size: the optimum
    equals BASELINE.md's KAT -520.38067984 (strong duality with the moment relaxation) and the oracle's dual run
    (tests/golden/make_golden_large.py bqpdual60), all residues <= 1e-8"""
    import json
    from manisdp_matlab_b200 import ManiDSDP_unitdiag
    gold = json.load(open(os.path.join(GOLDEN, "oracle_outputs_large.json")))
    A2, b, c, K, dAAt, maxb, _ = _sos(60)
    assert (K["s"], A2.shape[0]) == (gold["bqp_60_1_dual"]["n"], gold["bqp_60_1_dual"]["m"])
    assert (K["s"], A2.shape[0] - 1) == (1831, 523685)  # data/bqp_result.txt:27 (the authors' dual run at d = 60: 20.5 s)
    _, obj, data = ManiDSDP_unitdiag(A2, b, c, K, dict(dAAt=dAAt, tol=1e-8, line_search=1, verbose=False))
    assert data["status"] == 0 and max(data["gap"], data["pinf"], data["dinf"]) < 1e-8
    assert abs(obj * maxb - gold["bqp_60_1_dual"]["obj"]) <= 1e-6 * abs(gold["bqp_60_1_dual"]["obj"])
    assert abs(obj * maxb - gold["bqp_60_1_opt"]["obj"]) <= 1e-6 * abs(gold["bqp_60_1_opt"]["obj"])


def test_dual_closures_at_widths_beyond_512():
    """p = 640 on the SOS form of BQP-20: the dual closures (Gram terms included) beyond the 512-column row geometry"""
    from oracle.manisdp_ref import DualProblem
    A2, b, c, K, dAAt, _, _ = _sos(20)
    n, nf, m = K["s"], K["f"], A2.shape[0]
    A, B = A2[:, nf:].tocsr(), A2[:, :nf].tocsr()
    rng = np.random.default_rng(77)
    x, w = _state(n, m, nf, rng)
    x = _project_out(A, dAAt, x)
    sigma, p = 0.21, 640
    Y = rng.standard_normal((n, p))
    Y /= np.linalg.norm(Y, axis=1, keepdims=True)
    ora = DualProblem(A, B, b, c[nf:], c[:nf], n, p, dAAt, x, w, sigma)
    f0 = ora.cost(Y)
    g0 = ora.grad(Y)
    U = ora.M.proj(Y, rng.standard_normal((n, p)))
    H0 = ora.hess(Y, U)
    with _handle(A2, b, c, K, dAAt) as h:
        h.set_sigma(sigma)
        h.dual_set_state(x, w)
        h.set_Y(Y)
        f = h.cost()
        G, gn = h.grad()
        Hd = h.hess(U)
    assert abs(f - f0) <= 1e-11 * max(1.0, abs(f0))
    assert _rel(G, g0) < 1e-11 and _rel(Hd, H0) < 1e-10
