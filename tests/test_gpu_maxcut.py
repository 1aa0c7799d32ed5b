"""GPU parity tests of the ONLYUNITDIAG path through the C ABI.

The oracle (oracle/, NumPy restatement of the reference) is the checker; tolerances are FP64 reduction-order only:
1e-12 relative for single closure calls, 1e-7 relative on trust-region iterate logs (they compound over iterations),
1e-6 relative on optima as BASELINE.json's north_star states.
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gset(name):
    from manisdp_matlab_b200 import problems as P
    d = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    return P.maxcut_C(int(d["n"]), d["ei"].astype(np.int64), d["ej"].astype(np.int64), d["w"].astype(np.float64))


def _rand_point(n, p, seed):
    rng = np.random.default_rng(seed)
    Y = rng.standard_normal((n, p))
    return Y / np.linalg.norm(Y, axis=1, keepdims=True), rng


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(1e-300, np.linalg.norm(b))


@pytest.fixture(scope="module")
def G1():
    return _gset("G1")


@pytest.mark.parametrize("p", [2, 3, 7, 12, 40, 64, 100, 130, 300])
def test_closures_match_oracle(G1, p):
    """cost / grad / hess / proj / retr against the restated closures (ManiSDP_onlyunitdiag.m:117-156)."""
    from manisdp_matlab_b200 import Handle
    from oracle.manisdp_ref import OnlyUnitDiagProblem
    n = G1.shape[0]
    Y, rng = _rand_point(n, p, p)
    U = rng.standard_normal((n, p))
    prob = OnlyUnitDiagProblem(G1, p, stale_eG=False)
    f0 = prob.cost(Y)
    g0 = prob.grad(Y)
    with Handle("onlyunitdiag", n, C_csc=G1) as h:
        h.set_Y(Y)
        assert np.array_equal(h.get_Y(), Y)  # transfers are bit-exact
        f = h.cost()
        assert abs(f - f0) <= 1e-12 * abs(f0)
        g, gn = h.grad()
        assert _rel(g, g0) < 1e-12
        assert abs(gn - np.linalg.norm(g0)) <= 1e-12 * np.linalg.norm(g0)
        Ut = prob.M.proj(Y, U)
        assert _rel(h.project(U), Ut) < 1e-13
        assert _rel(h.hess(Ut), prob.hess(Y, Ut)) < 1e-12
        assert _rel(h.hess(U), prob.hess(Y, U)) < 1e-12  # also off the tangent space
        assert _rel(h.retract(0.3 * Ut), prob.M.retr(Y, 0.3 * Ut)) < 1e-13


@pytest.mark.parametrize("p", [2, 5, 8, 12, 16, 29, 32])
def test_narrow_row_kernel_matches_generic(G1, p, monkeypatch):
    """k_spmm_narrow (ld <= 32, the default there) and the generic k_spmm (MANISDP_SPMM_NARROW=0) do the same
    arithmetic in the same entry order: cost, gradient and Hessian product agree to rounding of the row dot product."""
    from manisdp_matlab_b200 import Handle
    n = G1.shape[0]
    Y, rng = _rand_point(n, p, 100 + p)
    U = rng.standard_normal((n, p))
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("MANISDP_SPMM_NARROW", flag)
        with Handle("onlyunitdiag", n, C_csc=G1) as h:
            h.set_Y(Y)
            f = h.cost()
            g, gn = h.grad()
            out[flag] = (f, g.copy(), gn, h.hess(U).copy())
    assert abs(out["1"][0] - out["0"][0]) <= 1e-14 * abs(out["0"][0])
    assert _rel(out["1"][1], out["0"][1]) < 1e-14
    assert abs(out["1"][2] - out["0"][2]) <= 1e-14 * out["0"][2]
    assert _rel(out["1"][3], out["0"][3]) < 1e-14


@pytest.mark.parametrize("p", [33, 40, 50, 64])
def test_block_major_product_matches_row_kernel(G1, p, monkeypatch):
    """MANISDP_SPMM_BM=2 forces the block-major entry stream (per-block partial rows + summing epilogue pass, the
    large-graph product) on a small instance: closures against the oracle and against the row kernel, and a
    trust-region solve (CUDA-graph and stream mode) with the same accept pattern and iterate."""
    from manisdp_matlab_b200 import Handle
    from oracle.manisdp_ref import OnlyUnitDiagProblem
    n = G1.shape[0]
    Y, rng = _rand_point(n, p, 200 + p)
    U = rng.standard_normal((n, p))
    prob = OnlyUnitDiagProblem(G1, p, stale_eG=False)
    out = {}
    for flag in ("2", "0"):
        monkeypatch.setenv("MANISDP_SPMM_BM", flag)
        with Handle("onlyunitdiag", n, C_csc=G1) as h:
            h.set_Y(Y)
            f = h.cost()
            g, gn = h.grad()
            hv = h.hess(U).copy()
            logs = []
            for use_graph in (1, 0):
                h.set_Y(Y)
                info = h.tr_solve(maxiter=6, maxinner=25, tolgradnorm=1e-9, use_graph=use_graph)
                logs.append(([(r.numinner, r.accepted, r.stop_inner) for r in h.tr_log()], info.cost, h.get_Y().copy()))
            out[flag] = (f, g.copy(), gn, hv, logs)
    f, g, gn, hv, logs = out["2"]
    assert abs(f - prob.cost(Y)) <= 1e-12 * abs(f)
    assert _rel(g, prob.grad(Y)) < 1e-12
    assert _rel(hv, prob.hess(Y, U)) < 1e-12
    assert _rel(hv, out["0"][3]) < 1e-13 and _rel(g, out["0"][1]) < 1e-13
    for (pat, cost, Yend), (pat0, cost0, Yend0) in zip(logs, out["0"][4]):
        assert abs(cost - cost0) <= 1e-6 * abs(cost0)  # (the two products differ in the last bits: no pattern check)
    assert logs[0][0] == logs[1][0] and _rel(logs[0][2], logs[1][2]) < 1e-9  # graph mode == stream mode


@pytest.mark.parametrize("p", [34, 40, 50, 64])
def test_lowdeg_batched_kernel_matches_row_kernel(p, monkeypatch):
    """k_spmm_lowdeg (32 rows per warp, staged (col, val) pairs; default on the toroidal profile for 32 < ld <= 64) does
    the row kernel's arithmetic in the row kernel's entry order: gradient and Hessian-product rows agree with
    MANISDP_SPMM_LOWDEG=0 to 1e-14 (rounding of the row sums), and everything matches the oracle (G11: 800-vertex torus, degree 4,
    n not a multiple of 32*8 so the ragged last batch is exercised)."""
    from manisdp_matlab_b200 import Handle
    from oracle.manisdp_ref import OnlyUnitDiagProblem
    G11 = _gset("G11")
    n = G11.shape[0]
    Y, rng = _rand_point(n, p, 300 + p)
    U = rng.standard_normal((n, p))
    prob = OnlyUnitDiagProblem(G11, p, stale_eG=False)
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("MANISDP_SPMM_LOWDEG", flag)
        with Handle("onlyunitdiag", n, C_csc=G11) as h:
            h.set_Y(Y)
            f = h.cost()
            g, gn = h.grad()
            hv = h.hess(U).copy()
            info = h.tr_solve(maxiter=5, maxinner=20, tolgradnorm=1e-9)
            k = h.kkt(delta=4, eig_tol=1e-9)
            out[flag] = (f, g.copy(), gn, hv, info.cost, h.get_Y().copy(), k.lam_min)
    f, g, gn, hv, cost, Yend, lam = out["1"]
    # same entry order per row; the row kernel keeps four partial accumulators per row, this one a single chain
    assert _rel(g, out["0"][1]) < 1e-14 and _rel(hv, out["0"][3]) < 1e-14
    assert abs(f - out["0"][0]) <= 1e-14 * abs(f) and abs(gn - out["0"][2]) <= 1e-14 * gn
    assert abs(cost - out["0"][4]) <= 1e-9 * abs(cost) and _rel(Yend, out["0"][5]) < 1e-7
    assert abs(lam - out["0"][6]) <= 1e-9 * (1 + abs(lam))
    assert abs(f - prob.cost(Y)) <= 1e-12 * abs(f)
    assert _rel(g, prob.grad(Y)) < 1e-12
    assert _rel(hv, prob.hess(Y, U)) < 1e-12


def test_nonsymmetric_and_empty_rows():
    """column lists of C are used as row lists: (Y*C)(:,j) = sum_i C(i,j) Y(:,i) also for a non-symmetric C with
    empty columns and an isolated vertex (ragged input)."""
    from manisdp_matlab_b200 import Handle
    rng = np.random.default_rng(5)
    n, p = 257, 6
    C = sp.random(n, n, density=0.02, random_state=7, format="lil")
    C[:, 17] = 0
    C[17, :] = 0
    C[:, 200] = 0
    C = sp.csc_matrix(C)
    Y, _ = _rand_point(n, p, 3)
    U = rng.standard_normal((n, p))
    YC = C.T @ Y
    eG = np.sum(YC * Y, axis=1, keepdims=True)
    with Handle("onlyunitdiag", n, C_csc=C) as h:
        h.set_Y(Y)
        assert abs(h.cost() - 0.5 * eG.sum()) < 1e-12 * abs(eG).sum()
        g, _ = h.grad()
        assert _rel(g, YC - Y * eG) < 1e-12
        eH = C.T @ U
        assert _rel(h.hess(U), eH - Y * np.sum(Y * eH, axis=1, keepdims=True) - U * eG) < 1e-12


def test_hessian_is_symmetric_and_matches_finite_differences(G1):
    """Manopt's checkhessian methodology: <U, H[V]> = <V, H[U]> on the tangent space and the second-order
    Taylor model along the retraction has slope 3."""
    from manisdp_matlab_b200 import Handle
    n, p = G1.shape[0], 10
    Y, rng = _rand_point(n, p, 11)
    with Handle("onlyunitdiag", n, C_csc=G1) as h:
        h.set_Y(Y)
        f0 = h.cost()
        g, _ = h.grad()
        U = h.project(rng.standard_normal((n, p)))
        V = h.project(rng.standard_normal((n, p)))
        HU, HV = h.hess(U), h.hess(V)
        assert abs(np.vdot(U, HV) - np.vdot(V, HU)) < 1e-10 * abs(np.vdot(U, HV))
        U *= np.sqrt(n) / np.linalg.norm(U)
        HU = h.hess(U)
        errs = []
        for t in [1e-1, 1e-2]:
            Yt = h.retract(t * U)
            h.set_Y(Yt)
            ft = h.cost()
            h.set_Y(Y)
            h.cost()
            errs.append(abs(ft - (f0 + t * np.vdot(g, U) + 0.5 * t * t * np.vdot(U, HU))))
        slope = np.log10(errs[0] / errs[1])
        assert 2.7 < slope < 3.3, (errs, slope)


@pytest.mark.parametrize("use_graph", [0, 1])
def test_tr_iterates_match_oracle_log(G1, use_graph):
    """Device RTR (graph WHILE loop and plain stream launches) reproduces the oracle's trust-region log from the same
    start: accept/reject pattern, inner iteration counts and tCG stop reasons identical; cost, gradnorm, rho to 1e-7."""
    from manisdp_matlab_b200 import Handle
    gold = json.load(open(os.path.join(GOLDEN, "oracle_outputs.json")))["G1_tr_log_seed123_p12"]
    n = G1.shape[0]
    Y0, _ = _rand_point(n, 12, 123)
    with Handle("onlyunitdiag", n, C_csc=G1) as h:
        h.set_Y(Y0)
        info = h.tr_solve(maxiter=12, maxinner=30, tolgradnorm=1e-8, use_graph=use_graph)
        log = h.tr_log()
    assert len(log) == len(gold)
    assert info.hv_count == sum(r["numinner"] for r in gold)
    for a, b in zip(log, gold):
        assert a.iter == b["iter"]
        assert a.accepted == int(b["accepted"])
        assert a.numinner == b["numinner"]
        assert a.stop_inner == b["stop_inner"]
        assert abs(a.cost - b["cost"]) <= 1e-9 * abs(b["cost"])
        assert abs(a.gradnorm - b["gradnorm"]) <= 1e-6 * max(b["gradnorm"], 1e-6)
        assert abs(a.Delta - b["Delta"]) <= 1e-12 * b["Delta"]
        if b["rho"] is not None:
            assert abs(a.rho - b["rho"]) <= 1e-6 * max(1.0, abs(b["rho"]))


def test_tr_live_oracle_other_widths(G1):
    """same comparison against the oracle run live, on widths that exercise other row-group geometries"""
    from manisdp_matlab_b200 import Handle
    from oracle.manisdp_ref import OnlyUnitDiagProblem
    from oracle.manopt_rtr import trustregions
    n = G1.shape[0]
    for p in [5, 70]:
        Y0, _ = _rand_point(n, p, 1000 + p)
        res = trustregions(OnlyUnitDiagProblem(G1, p, stale_eG=False), Y0.copy(), maxiter=6, maxinner=25,
                           tolgradnorm=1e-8)
        with Handle("onlyunitdiag", n, C_csc=G1) as h:
            h.set_Y(Y0)
            h.tr_solve(maxiter=6, maxinner=25, tolgradnorm=1e-8)
            log = h.tr_log()
            Yd = h.get_Y()
        assert [r.numinner for r in log] == [r.numinner for r in res.info]
        assert [r.accepted for r in log] == [int(r.accepted) for r in res.info]
        assert abs(log[-1].cost - res.cost) <= 1e-9 * abs(res.cost)
        assert _rel(Yd, res.x) < 1e-6


def test_eig_step_matches_dense_eig(G1):
    """device LOBPCG on S = C - diag(z) against numpy's full eig (the reference's eig(full(S)), :50)."""
    from manisdp_matlab_b200 import Handle
    n = G1.shape[0]
    Y0, _ = _rand_point(n, 8, 77)
    with Handle("onlyunitdiag", n, C_csc=G1) as h:
        h.set_Y(Y0)
        h.tr_solve(maxiter=3, maxinner=20, tolgradnorm=1e-8)
        k = h.kkt(8, 1e-10, 0)
        vals, vecs = h.get_eigs(8)
        Y = h.get_Y()
    X = Y @ Y.T
    z = np.asarray(G1.multiply(X).sum(axis=0)).ravel()
    S = G1.toarray() - np.diag(z)
    dS, vS = np.linalg.eigh(S)
    assert abs(k.obj - z.sum()) <= 1e-10 * abs(z.sum())
    assert np.allclose(vals, dS[:8], atol=1e-7 * (1 + abs(dS[-1])))
    assert abs(k.lam_max - dS[-1]) <= 1e-3 * abs(dS[-1])
    dinf = max(0.0, -dS[0]) / (1 + dS[-1])
    assert abs(k.dinf - dinf) <= 1e-3 * dinf + 1e-9
    assert k.nneg == min(int((dS < 0).sum()), 8)
    # returned vectors are eigenvectors: residual small relative to the spectrum width
    R = S @ vecs - vecs * vals
    assert np.linalg.norm(R, axis=0).max() < 1e-6 * (1 + abs(dS[-1]))


@pytest.mark.parametrize("delta", [13, 16, 24])
def test_eig_step_wide_block(G1, delta):
    """options.delta > 12 makes the LOBPCG basis wider than the 48-column register Gram (round-1 advisor finding): the
    tiled Gram path must deliver the same `delta` smallest eigenpairs as a dense eigh; delta > 60 is rejected loudly."""
    from manisdp_matlab_b200 import Handle
    from manisdp_matlab_b200._lib import EngineError
    n = G1.shape[0]
    Y0, _ = _rand_point(n, 10, 78)
    with Handle("onlyunitdiag", n, C_csc=G1) as h:
        h.set_Y(Y0)
        h.tr_solve(maxiter=3, maxinner=20, tolgradnorm=1e-8)
        k = h.kkt(delta, 1e-9, 0)
        vals, vecs = h.get_eigs(delta)
        Y = h.get_Y()
        with pytest.raises(EngineError):
            h.kkt(61, 1e-10, 0)
    X = Y @ Y.T
    z = np.asarray(G1.multiply(X).sum(axis=0)).ravel()
    S = G1.toarray() - np.diag(z)
    dS, _ = np.linalg.eigh(S)
    assert np.allclose(vals, dS[:delta], atol=1e-7 * (1 + abs(dS[-1]))), (vals - dS[:delta], k.eig_resid, k.eig_iters)
    assert k.nneg == min(int((dS < 0).sum()), delta)
    R = S @ vecs - vecs * vals
    assert np.linalg.norm(R, axis=0).max() < 1e-6 * (1 + abs(dS[-1]))
    # eig_converged reports the strict residual test; a block that stopped on stagnation just above it (clustered wanted
    # values) must still be within a decade of the tolerance
    assert k.eig_converged == 1 or k.eig_resid <= 1e-8 * (1 + abs(dS[-1])), (k.eig_resid, k.eig_iters)


def test_rank_cut_matches_svd(G1):
    from manisdp_matlab_b200 import Handle
    n = G1.shape[0]
    rng = np.random.default_rng(9)
    B = rng.standard_normal((n, 4))
    Y = np.hstack([B, B @ rng.standard_normal((4, 5)) * 1e-6])  # numerical rank 4, width 9
    Y /= np.linalg.norm(Y, axis=1, keepdims=True)
    e = np.linalg.svd(Y, compute_uv=False)
    r0 = int((e >= 1e-3 * e[0]).sum())
    with Handle("onlyunitdiag", n, C_csc=G1) as h:
        h.set_Y(Y)
        r, pn = h.rank_cut(1e-3, apply=True)
        Yc = h.get_Y()
    assert r == r0 == 4 and pn == 4
    Us, es, _ = np.linalg.svd(Y, full_matrices=False)
    assert _rel(Yc @ Yc.T, (Us[:, :4] * es[:4]) @ (Us[:, :4] * es[:4]).T) < 1e-10


def test_escape_appends_and_normalises(G1):
    from manisdp_matlab_b200 import Handle
    n = G1.shape[0]
    Y0, _ = _rand_point(n, 6, 4)
    with Handle("onlyunitdiag", n, C_csc=G1) as h:
        h.set_Y(Y0)
        h.tr_solve(maxiter=2, maxinner=10, tolgradnorm=1e-8)
        h.kkt(8, 1e-8, 0)
        vals, vecs = h.get_eigs(3)
        Y = h.get_Y()
        h.escape(3, 0.5, 0)
        Yn = h.get_Y()
    ref = np.hstack([Y, 0.5 * vecs])
    ref /= np.linalg.norm(ref, axis=1, keepdims=True)
    assert Yn.shape == (n, 9)
    assert _rel(Yn, ref) < 1e-13


@pytest.mark.parametrize("name,p0", [("G1", 40), ("G11", 40)])
def test_full_solve_reaches_kkt_and_known_optimum(name, p0):
    """BASELINE config 1: optimum to rel. 1e-6 of the oracle / SDPLIB value, dinf <= tol = 1e-8."""
    from manisdp_matlab_b200 import ManiSDP_onlyunitdiag
    gold = json.load(open(os.path.join(GOLDEN, "oracle_outputs.json")))[f"{name}_opt"]
    C = _gset(name)
    X, obj, data = ManiSDP_onlyunitdiag(C, dict(p0=p0, verbose=False))
    assert data["status"] == 0
    assert data["dinf"] < 1e-8
    assert abs(obj - gold["obj"]) <= 1e-6 * abs(gold["obj"])
    # independent check of the reported residue with a dense eig on the host
    dS = np.linalg.eigvalsh(data["S"])
    assert max(0.0, -dS[0]) / (1 + dS[-1]) < 1e-7
    assert np.allclose(np.diag(X), 1.0, atol=1e-12)


def test_known_answer_G1_literal():
    from manisdp_matlab_b200 import ManiSDP_onlyunitdiag
    X, obj, data = ManiSDP_onlyunitdiag(_gset("G1"), dict(p0=40, verbose=False))
    assert abs(obj - (-12083.19765455)) <= 1e-6 * 12083.2  # SURVEY 8c KAT


def test_device_rand_is_reproducible_and_on_manifold(G1):
    from manisdp_matlab_b200 import Handle
    n = G1.shape[0]
    with Handle("onlyunitdiag", n, C_csc=G1) as h:
        h.rand_Y(9, 42)
        A = h.get_Y()
        h.rand_Y(9, 42)
        B = h.get_Y()
        h.rand_Y(9, 43)
        Cc = h.get_Y()
    assert np.array_equal(A, B) and not np.array_equal(A, Cc)
    assert np.allclose(np.linalg.norm(A, axis=1), 1.0, atol=1e-14)
    assert abs(A.mean()) < 0.02
