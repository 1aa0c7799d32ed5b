"""CPU: pin the oracle (oracle/, the NumPy restatement of the reference) against known answers that do not come from
the oracle itself -- SDPLIB optimal values shipped in the reference tree (data/sdplib/README:71), an exhaustive BQP
minimum, a DIMACS theta value -- and against the committed golden outputs (tests/golden/oracle_outputs.json).
The reference ships no expected optima of its own and cannot run here (no MATLAB): SURVEY 8c."""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(GOLDEN, "oracle_outputs.json")))


def _gset(name):
    from oracle import generators as g
    d = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    return g.maxcut_C(int(d["n"]), d["ei"].astype(np.int64), d["ej"].astype(np.int64), d["w"].astype(np.float64))


def test_get_basis_matches_sequential_rule():
    from oracle import generators as g
    for n, d in [(1, 3), (2, 2), (3, 2), (4, 2), (3, 4), (6, 2)]:
        assert np.array_equal(g.get_basis(n, d), g.get_basis_sequential(n, d))


def test_maxcut_G11_matches_sdplib(gold):
    """SDPLIB maxG11 optimal value 629.1648 (reference data/sdplib/README:71)."""
    from oracle import manisdp_ref as ref
    X, obj, data = ref.ManiSDP_onlyunitdiag(_gset("G11"), dict(p0=40, seed=0))
    assert data["dinf"] < 1e-8
    assert abs(-obj - 629.1648) < 5e-4
    assert abs(obj - gold["G11_opt"]["obj"]) <= 1e-8 * abs(obj)


def test_maxcut_G1_known_answer(gold):
    from oracle import manisdp_ref as ref
    X, obj, data = ref.ManiSDP_onlyunitdiag(_gset("G1"), dict(p0=40, seed=0))
    assert abs(obj - (-12083.19765455)) < 1e-6 * 12083.2
    assert abs(obj - gold["G1_opt"]["obj"]) <= 1e-9 * abs(obj)
    assert np.allclose(np.diag(X), 1.0)


def test_bqp10_matches_bruteforce(gold):
    from oracle import generators as g
    from oracle import manisdp_ref as ref
    d = np.load(os.path.join(GOLDEN, "bqp_10_1.npz"))
    At, b, c, K = g.bqpmom(10, d["Q"], d["e"])
    assert (int(K["s"]), At.shape[1]) == (66 - 10, 1256)  # n = 1 + 10 + 45, m as in data/bqp_result.txt
    mc = float(np.abs(c).max())
    X, obj, data = ref.ManiSDP_unitdiag(At, b, c / mc, K, dict(seed=0))
    brute = g.bqp_bruteforce(d["Q"], d["e"])
    assert abs(obj * mc - brute) <= 1e-7 * abs(brute)
    assert abs(brute - gold["bqp_10_1_bruteforce"]) < 1e-12
    assert max(data["gap"], data["pinf"], data["dinf"]) < 1e-8


def test_theta_hamming_7_5_6_matches_dimacs(gold):
    from oracle import generators as g
    from oracle import manisdp_ref as ref
    At, b, c, K = g.generate_hamming(7, [5, 6])
    assert int(K["s"]) == 128
    X, obj, data = ref.ManiSDP_unittrace(At, b, c, K, dict(seed=0, tol=1e-6, sigma0=1e5, sigma_max=1e8, line_search=1))
    assert abs(-obj - 128.0 / 3.0) < 1e-4  # DIMACS hamming_7_5_6: 42.6667
    assert abs(obj - gold["hamming_7_5_6_opt"]["obj"]) < 1e-5


def test_qsphere10_golden(gold):
    from oracle import generators as g
    from oracle import manisdp_ref as ref
    d = np.load(os.path.join(GOLDEN, "qs_c_10_1.npz"))
    At, b, c, K = g.qsmom(10, d["coe"])
    X, obj, data = ref.ManiSDP(At, b, c, K, dict(seed=0, tol=1e-8, theta=1e-2, tau1=0.02))
    assert abs(obj - (-5.34235276)) < 1e-7
    assert abs(obj - gold["qs_c_10_1_opt"]["obj"]) < 1e-8


def test_tr_log_golden_is_reproducible(gold):
    """the committed trust-region log is what the oracle produces today (guards the fixture against drift)"""
    from oracle.manisdp_ref import OnlyUnitDiagProblem
    from oracle.manopt_rtr import trustregions
    C = _gset("G1")
    rng = np.random.default_rng(123)
    Y0 = rng.standard_normal((C.shape[0], 12))
    Y0 /= np.linalg.norm(Y0, axis=1, keepdims=True)
    res = trustregions(OnlyUnitDiagProblem(C, 12, stale_eG=False), Y0, maxiter=12, maxinner=30, tolgradnorm=1e-8)
    log = gold["G1_tr_log_seed123_p12"]
    assert [r.numinner for r in res.info] == [r["numinner"] for r in log]
    assert np.allclose([r.cost for r in res.info], [r["cost"] for r in log], rtol=1e-10)


def test_product_laplacian_matches_oracle_and_reference_semantics():
    """problems.laplacian (sparse) == oracle restatement of Laplacian.m, incl. the duplicate-edge rule (:7 assigns,
    :9-10 accumulate)."""
    from manisdp_matlab_b200 import problems as P
    from oracle import generators as g
    ei = np.array([0, 1, 1, 2, 0]); ej = np.array([1, 2, 0, 3, 1]); w = np.array([1.0, 2.0, 5.0, -1.0, 7.0])
    A = P.laplacian(4, ei, ej, w).toarray()
    B = g.laplacian(4, ei, ej, w).toarray()
    assert np.array_equal(A, B)
    assert A[0, 1] == -7.0 and A[0, 0] == 1.0 + 5.0 + 7.0


def test_multiblock_oracle_pinned_by_the_single_block_driver():
    """ManiSDP_multiblock's restatement has no optimum of its own in the reference tree; it is pinned here through the
    block-diagonal embedding: with K.nob = 0 the multi-block optimum equals the optimum of ONE PSD cone of order
    sum(n_i) whose off-diagonal blocks are free, solved by the ManiSDP.m restatement (itself pinned on SDPLIB above)."""
    from instances import generators as G
    from oracle.manisdp_ref import ManiSDP, ManiSDP_multiblock
    At, b, c, K = G.multiblock_random([6, 4, 5, 3], 0, 4, 1)
    _, o1, d1 = ManiSDP_multiblock(At, b, c, K, dict(tol=1e-8))
    Ab, cb, N, _ = G.embed_multiblock(At, c, K)
    _, o2, d2 = ManiSDP(Ab, b, cb, {"s": N}, dict(tol=1e-8))
    assert d1["status"] == 0 and d2["status"] == 0
    assert abs(o1 - o2) <= 1e-7 * max(1.0, abs(o2))
    # unit-diagonal blocks: feasibility and a dual certificate from the returned multipliers
    At, b, c, K = G.multiblock_random([8, 6, 7], 3, 6, 3)
    X, o3, d3 = ManiSDP_multiblock(At, b, c, K, dict(tol=1e-8))
    assert d3["status"] == 0
    for Xi in X:
        assert np.allclose(np.diag(Xi), 1.0, atol=1e-12)
    assert all(np.linalg.eigvalsh(S)[0] > -1e-6 for S in d3["S"])  # dual feasibility of the slack blocks


def test_dual_oracle_pinned_by_bruteforce_and_strong_duality():
    """ManiDSDP_unitdiag's restatement (src/dual/ManiDSDP_unitdiag.m) on the SOS form of the reference's BQP data files:
    d = 10 reaches the exhaustive minimum over {-1,+1}^10 (the relaxation is tight there), d = 20 the optimum of the
    primal moment relaxation already pinned above (strong duality), with the options of example/dual/example_bqp_dual.m."""
    import scipy.sparse as sp
    from instances import generators as G
    from oracle.manisdp_ref import ManiDSDP_unitdiag
    gold = json.load(open(os.path.join(GOLDEN, "oracle_outputs.json")))
    for q, target in [(10, None), (20, gold["bqp_20_1_opt"]["obj"])]:
        d = np.load(os.path.join(GOLDEN, f"bqp_{q}_1.npz"))
        A, b, dAAt, mb = G.bqpsos(d["Q"], d["e"], q)
        v = np.zeros((A.shape[0], 1))
        v[0] = 1.0
        A2 = sp.hstack([sp.csr_matrix(v), A]).tocsr()
        c = np.concatenate([[1.0], np.zeros(mb * mb)])
        maxb = float(np.abs(b).max())
        X, obj, data = ManiDSDP_unitdiag(A2, b / maxb, c, {"f": 1, "s": mb}, dict(dAAt=dAAt, tol=1e-8, line_search=1))
        assert data["status"] == 0 and max(data["gap"], data["pinf"], data["dinf"]) < 1e-8
        if target is None:
            target = G.bqp_bruteforce(d["Q"], d["e"])
        assert abs(obj * maxb - target) <= 1e-6 * abs(target), (q, obj * maxb, target)
        assert np.allclose(np.diag(data["S"]), 1.0, atol=1e-12) and np.linalg.eigvalsh(X)[0] > -1e-6
