"""GPU parity tests of the multi-block driver (src/primal/ManiSDP_multiblock.m:60-249,
src/basicfunction/multiblockmanifold.m:1-42, src/C-files/{projc,retrc,innerc,lincombc,randc}.cpp) through the C ABI
(manisdp_mb_* + the shared closure / solver entry points), against oracle/manisdp_ref.py::ManiSDP_multiblock."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# (block orders, nob, m, seed) of instances.generators.multiblock_random on which the oracle reaches 1e-8
CASES = [([6, 4, 5, 3], 0, 4, 1), ([6, 4, 5, 3], 2, 4, 2), ([8, 6, 7], 3, 6, 3), ([8, 6, 7, 2], 2, 5, 4)]


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(1e-300, np.linalg.norm(b))


def _instance(case):
    from instances import generators as G
    ns, nob, m, seed = case
    return G.multiblock_random(ns, nob, m, seed)


def _point(ns, nob, widths, rng):
    Y = []
    for i, (n, p) in enumerate(zip(ns, widths)):
        B = rng.standard_normal((n, p))
        if i < nob:
            B /= np.linalg.norm(B, axis=1, keepdims=True)
        Y.append(B)
    return Y


def _handle(At, b, c, K, **kw):
    from manisdp_matlab_b200 import Handle
    ns = [int(v) for v in K["s"]]
    return Handle("multiblock", int(sum(ns)), At=At, b=b, c=c, block_sizes=ns, nob=int(K["nob"]), **kw)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("widths", ["ones", "ragged"])
def test_multiblock_closures_match_oracle(case, widths):
    """cost / grad / hess of ManiSDP_multiblock.m:203-247 and the manifold operations at a random point whose blocks
    have different widths, with a non-trivial dual vector and penalty."""
    from oracle.manisdp_ref import MultiblockProblem
    from oracle.manopt_rtr import Cells
    At, b, c, K = _instance(case)
    ns, nob = K["s"], K["nob"]
    rng = np.random.default_rng(17)
    p = [1] * len(ns) if widths == "ones" else [min(n, 1 + (3 * i + 2) % 5) for i, n in enumerate(ns)]
    Y = _point(ns, nob, p, rng)
    y = 0.3 * rng.standard_normal(At.shape[1])
    sigma = 0.7
    ora = MultiblockProblem(At.tocsc(), b, c, ns, p, nob, y, sigma)
    Yc = Cells(Y)
    f0 = ora.cost(Yc)
    g0 = ora.grad(Yc)
    U = ora.M.proj(Yc, Cells([rng.standard_normal(B.shape) for B in Y]))
    H0 = ora.hess(Yc, U)
    E = Cells([0.3 * rng.standard_normal(B.shape) for B in Y])
    R0 = ora.M.retr(Yc, E)
    with _handle(At, b, c, K) as h:
        st = h.stats()
        assert (st.s_mode, st.a_mode) == (1, 1)  # sparse representation: nothing of size N x N
        h.set_dual(y, sigma)
        h.mb_set_Y(Y)
        assert h.mb_widths() == p
        for Ba, Bb in zip(h.mb_get_Y(), Y):
            assert np.array_equal(Ba, Bb)
        f = h.cost()
        G, gn = h.grad()
        Hd = h.hess(h.mb_join(U.b))
        Pd = h.project(h.mb_join(E.b))
        Rd = h.retract(h.mb_join(E.b))
        # the columns beyond p_i stay exactly zero
        full = np.zeros((h.n, h.p), dtype=bool)
        r = 0
        for n, q in zip(ns, p):
            full[r:r + n, :q] = True
            r += n
        for A in (G, Hd, Pd, Rd):
            assert np.all(A[~full] == 0.0)
        Gs, Hs, Ps, Rs = h.mb_split(G), h.mb_split(Hd), h.mb_split(Pd), h.mb_split(Rd)
    assert abs(f - f0) <= 1e-12 * max(1.0, abs(f0))
    P0 = ora.M.proj(Yc, E)
    for i in range(len(ns)):
        assert _rel(Gs[i], g0[i]) < 1e-12
        assert _rel(Hs[i], H0[i]) < 1e-11
        assert _rel(Ps[i], P0[i]) < 1e-13
        assert _rel(Rs[i], R0[i]) < 1e-13
    assert abs(gn - ora.M.norm(Yc, g0)) <= 1e-12 * max(1.0, gn)


def test_multiblock_index_split_is_exact():
    """row r of At -> (block, i, j) -> (row, column) of the embedded matrix, read back as integers"""
    At, b, c, K = _instance(([5, 3, 4, 1, 2], 2, 9, 11))
    ns = K["s"]
    Atc = At.tocsc()
    Atc.sort_indices()
    off2 = np.concatenate([[0], np.cumsum([n * n for n in ns])])
    roff = np.concatenate([[0], np.cumsum(ns)])
    r = Atc.indices.astype(np.int64)
    blk = np.searchsorted(off2, r, side="right") - 1
    loc = r - off2[blk]
    nb = np.asarray(ns)[blk]
    with _handle(At, b, c, K) as h:
        i, j = h.index_split()
    assert np.array_equal(i, roff[blk] + loc % nb)
    assert np.array_equal(j, roff[blk] + loc // nb)


@pytest.mark.parametrize("case", CASES[1:3])
@pytest.mark.parametrize("use_graph", [1, 0])
def test_multiblock_tr_iterates_match_oracle(case, use_graph):
    """trustregions + tCG on the product manifold: accept pattern, inner counts and stop codes identical"""
    from oracle.manisdp_ref import MultiblockProblem
    from oracle.manopt_rtr import Cells, trustregions
    At, b, c, K = _instance(case)
    ns, nob = K["s"], K["nob"]
    rng = np.random.default_rng(5)
    p = [min(n, 2 + i % 3) for i, n in enumerate(ns)]
    Y0 = _point(ns, nob, p, rng)
    y = 0.1 * rng.standard_normal(At.shape[1])
    sigma = 2.0
    ora = MultiblockProblem(At.tocsc(), b, c, ns, p, nob, y, sigma)
    res = trustregions(ora, Cells([B.copy() for B in Y0]), maxiter=6, maxinner=15, tolgradnorm=1e-10)
    with _handle(At, b, c, K) as h:
        h.set_dual(y, sigma)
        h.mb_set_Y(Y0)
        info = h.tr_solve(maxiter=6, maxinner=15, tolgradnorm=1e-10, use_graph=use_graph)  # Delta_bar: M.typicaldist()
        log = h.tr_log()
        Yd = h.mb_get_Y()
    assert [r.numinner for r in log] == [r.numinner for r in res.info]
    assert [r.accepted for r in log] == [int(r.accepted) for r in res.info]
    assert [r.stop_inner for r in log] == [r.stop_inner for r in res.info]
    assert abs(log[0].Delta - ora.M.typicaldist() / 8) <= 1e-14 * ora.M.typicaldist()
    assert abs(info.cost - res.cost) <= 1e-9 * max(1.0, abs(res.cost))
    for i in range(len(ns)):
        assert _rel(Yd[i], res.x[i]) < 1e-7


@pytest.mark.parametrize("case", CASES[:3])
@pytest.mark.parametrize("eig", ["host", "device"])
def test_multiblock_kkt_matches_dense_formulas(case, eig, monkeypatch):
    """ManiSDP_multiblock.m:66-97 against a dense per-block NumPy evaluation (eig of every S{i}); both block eigensolvers:
    Householder + QL on host threads (default) and the batched device Jacobi (MANISDP_MB_EIG=device, csrc/jacobi.cu)"""
    monkeypatch.setenv("MANISDP_MB_EIG", eig)
    At, b, c, K = _instance(case)
    ns, nob = K["s"], K["nob"]
    rng = np.random.default_rng(3)
    p = [min(n, 2 + i % 2) for i, n in enumerate(ns)]
    Y = _point(ns, nob, p, rng)
    y = 0.05 * rng.standard_normal(At.shape[1])
    sigma = 3.0
    X = [B @ B.T for B in Y]
    x = np.concatenate([Xi.reshape(-1, order="F") for Xi in X])
    Axb = At.T @ x - b
    y1 = y - sigma * Axb
    cy = c - At @ y1
    by = float(b @ y1)
    off2 = np.concatenate([[0], np.cumsum([n * n for n in ns])])
    evs, dinfs, S = [], [], []
    for i, n in enumerate(ns):
        Si = cy[off2[i]:off2[i + 1]].reshape(n, n, order="F")
        if i < nob:
            z = np.sum(X[i] * Si, axis=0)
            by += float(z.sum())
            Si = Si - np.diag(z)
        d = np.linalg.eigvalsh(Si)
        S.append(Si)
        evs.append(d)
        dinfs.append(max(0.0, -d[0]) / (1 + abs(d[-1])))
    obj = float(c @ x)
    with _handle(At, b, c, K) as h:
        h.set_dual(y, sigma)
        h.mb_set_Y(Y)
        k, dd, nneg = h.mb_kkt(update_dual=1)
        y_dev, _ = h.get_dual()
        for i, n in enumerate(ns):
            vals, vecs = h.mb_block_eigs(i)
            assert np.allclose(vals, evs[i], rtol=0, atol=1e-11 * max(1.0, np.abs(evs[i]).max()))
            assert np.linalg.norm(S[i] @ vecs - vecs * vals) < 1e-10 * max(1.0, np.abs(evs[i]).max())
            assert np.linalg.norm(vecs.T @ vecs - np.eye(n)) < 1e-12
    assert abs(k.obj - obj) <= 1e-12 * max(1.0, abs(obj))
    assert abs(k.pinf - np.linalg.norm(Axb) / (1 + np.linalg.norm(b))) <= 1e-12
    assert abs(k.by - by) <= 1e-11 * max(1.0, abs(by))
    assert abs(k.gap - abs(obj - by) / (abs(by) + abs(obj) + 1)) <= 1e-12
    assert np.allclose(dd, dinfs, rtol=0, atol=1e-12)
    assert abs(k.dinf - max(dinfs)) <= 1e-12
    assert list(nneg) == [int(np.sum(d < 0)) for d in evs]
    assert _rel(y_dev, y1) < 1e-13


@pytest.mark.parametrize("line_search", [0, 1])
def test_multiblock_update_matches_reference_rules(line_search):
    """rank cut + escape of every block (ManiSDP_multiblock.m:114-153): widths, X_i = Y_i Y_i' of the new point (the
    truncation basis is unique up to signs, X is not affected) and, with line_search = 1, the accepted step."""
    from oracle.manisdp_ref import MultiblockProblem, _mb_line_search
    from oracle.manopt_rtr import Cells
    At, b, c, K = _instance(([8, 6, 7, 1, 5], 2, 5, 21))
    ns, nob = K["s"], K["nob"]
    rng = np.random.default_rng(8)
    # block 0: numerically rank 2 of width 4 (cut expected); block 2: width 1; block 3: order 1 (< min_facsize)
    p = [4, 3, 1, 1, 5]
    Y = _point(ns, nob, p, rng)
    Y[0][:, 2:] *= 1e-5
    Y[0] /= np.linalg.norm(Y[0], axis=1, keepdims=True)
    y = 0.2 * rng.standard_normal(At.shape[1])
    sigma, theta, delta, alpha = 1.5, 1e-2, 3, 0.1
    # the reference's rules on the host
    X = [B @ B.T for B in Y]
    x = np.concatenate([Xi.reshape(-1, order="F") for Xi in X])
    y1 = y - sigma * (At.T @ x - b)
    cy = c - At @ y1
    off2 = np.concatenate([[0], np.cumsum([n * n for n in ns])])
    Yn, Un, pn = [], [], []
    for i, n in enumerate(ns):
        Si = cy[off2[i]:off2[i + 1]].reshape(n, n, order="F")
        if i < nob:
            Si = Si - np.diag(np.sum(X[i] * Si, axis=0))
        d, v = np.linalg.eigh(Si)
        Yi, pi = Y[i], p[i]
        if n < 2:
            Yn.append(Yi); Un.append(np.zeros_like(Yi)); pn.append(pi)
            continue
        if pi > 1:
            Us, e, _ = np.linalg.svd(Yi, full_matrices=False)
            r = max(1, int(np.sum(e >= theta * e[0])))
            if r < pi:
                Yi, pi = Us[:, :r] * e[:r], r
        nneg = int(np.sum(d < 0))
        nne = max(min(nneg, delta), 1) if i < nob else min(nneg, delta)
        if pi + nne > n:
            nne = 0
        V = v[:, :nne]
        if line_search:
            Un.append(np.hstack([np.zeros((n, pi)), V]))
            Yi = np.hstack([Yi, np.zeros((n, nne))])
        else:
            Yi = np.hstack([Yi, alpha * V])
            if i < nob:
                Yi = Yi / np.linalg.norm(Yi, axis=1, keepdims=True)
        Yn.append(Yi)
        pn.append(pi + nne)
    assert pn[0] < p[0] + delta  # the cut of block 0 happened
    with _handle(At, b, c, K) as h:
        h.set_dual(y, sigma)
        h.mb_set_Y(Y)
        h.mb_kkt(update_dual=1)
        pd = h.mb_update(theta, delta, alpha, line_search=line_search, min_facsize=2)
        assert pd == pn and h.mb_widths() == pn
        Yd = h.mb_get_Y()
        for i in range(len(ns)):
            assert _rel(Yd[i] @ Yd[i].T, Yn[i] @ Yn[i].T) < 1e-10
        if line_search:
            Ud = h.mb_split(h.slot_get(7))
            for i in range(len(ns)):
                assert _rel(Ud[i] @ Ud[i].T, Un[i] @ Un[i].T) < 1e-10
                assert _rel(Yd[i] @ Ud[i].T, Yn[i] @ Un[i].T) < 1e-10 or np.linalg.norm(Yn[i] @ Un[i].T) < 1e-14
            sigma2 = 2 * sigma
            h.set_sigma(sigma2)
            a = h.line_search()
            Yl = h.mb_get_Y()
            ora = MultiblockProblem(At.tocsc(), b, c, ns, pn, nob, y1, sigma2)
            ref = _mb_line_search(ora.co, Cells(Yn), Cells(Un), nob)
            co_dev = ora.co(Cells(Yl))
            co_ref = ora.co(ref)
            assert abs(co_dev - co_ref) <= 1e-9 * max(1.0, abs(co_ref))
            assert 0.8 ** 15 * 0.999 <= a <= 1.0


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("line_search", [0, 1])
def test_multiblock_full_solve_matches_oracle(case, line_search):
    """ManiSDP_multiblock end to end: same optimum as the oracle's restatement, all residues <= tol"""
    from manisdp_matlab_b200 import ManiSDP_multiblock
    from oracle.manisdp_ref import ManiSDP_multiblock as ref_mb
    At, b, c, K = _instance(case)
    opts = dict(tol=1e-8, AL_maxiter=400, line_search=line_search, verbose=False)
    _, obj_ref, dref = ref_mb(At, b, c, K, opts)
    assert dref["status"] == 0
    X, obj, data = ManiSDP_multiblock(At, b, c, K, opts)
    assert data["status"] == 0
    assert max(data["gap"], data["pinf"], data["dinf"]) < 1e-8
    assert abs(obj - obj_ref) <= 1e-6 * max(1.0, abs(obj_ref))
    # feasibility of the returned blocks, evaluated on the host
    x = np.concatenate([Xi.reshape(-1, order="F") for Xi in X])
    assert np.linalg.norm(At.T @ x - b) / (1 + np.linalg.norm(b)) < 1e-7
    for i in range(K["nob"]):
        assert np.allclose(np.diag(X[i]), 1.0, atol=1e-12)


def test_multiblock_equals_single_block_general_on_the_embedding():
    """independent pin of the optimum: with K.nob = 0 the multi-block problem and its block-diagonal embedding into one
    PSD cone (solved by the ManiSDP.m driver, pinned on SDPLIB) have the same optimal value"""
    from instances import generators as G
    from manisdp_matlab_b200 import ManiSDP, ManiSDP_multiblock
    At, b, c, K = _instance(CASES[0])
    _, obj_mb, d1 = ManiSDP_multiblock(At, b, c, K, dict(tol=1e-8, verbose=False))
    Ab, cb, N, _ = G.embed_multiblock(At, c, K)
    _, obj_g, d2 = ManiSDP(Ab, b, cb, {"s": N}, dict(tol=1e-8, verbose=False))
    assert d1["status"] == 0 and d2["status"] == 0
    assert abs(obj_mb - obj_g) <= 1e-6 * max(1.0, abs(obj_g))


# ---- the reference's own multi-block example family: sparse BQP moment relaxations (example/example_bqp_sparse.m) -------
@pytest.mark.parametrize("t,q,seed", [(3, 4, 1), (3, 5, 2), (4, 4, 3)])
def test_sparse_bqp_relaxation_is_tight_on_small_instances(t, q, seed, monkeypatch):
    """oracle-free pin: bqpmom_sparse + ManiSDP_multiblock (all blocks unit-diagonal, options of
    example_bqp_sparse.m:25-29) reach the exhaustive minimum of the clique-sparse BQP over {-1,+1}^n"""
    from instances import generators as G
    from manisdp_matlab_b200 import ManiSDP_multiblock
    At, b, c, K, n, I, coe = G.bqp_sparse_instance(t, q, seed)
    if seed == 3:  # one of the three instances goes through the device block eigensolver
        monkeypatch.setenv("MANISDP_MB_EIG", "device")
    X, obj, data = ManiSDP_multiblock(At, b, c, K, dict(tol=1e-8, line_search=1, tau1=1, verbose=False))
    assert data["status"] == 0 and max(data["gap"], data["pinf"], data["dinf"]) < 1e-8
    assert abs(obj - G.bqp_sparse_bruteforce(n, I, coe)) <= 1e-6 * max(1.0, abs(obj))
    for Xi in X:
        assert np.allclose(np.diag(Xi), 1.0, atol=1e-12)


def test_sparse_bqp_example_at_its_stated_size():
    """example/example_bqp_sparse.m:4-29 at t = 20 cliques of q = 20 variables (20 blocks of order 211, m = 327 315,
    nnz(At) = 1.4 M): optimum pinned by the oracle (tests/golden/make_golden_large.py bqpsparse), residues <= 1e-8"""
    import json
    import os
    from instances import generators as G
    from manisdp_matlab_b200 import ManiSDP_multiblock
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                       "oracle_outputs_large.json")))["bqp_sparse_20_20"]
    At, b, c, K, n, I, coe = G.bqp_sparse_instance(20, 20, 1)
    assert (At.shape[1], At.nnz, K["s"][0]) == (gold["m"], gold["nnz"], 211)
    # the outer iteration of this example is chaotic in the rounding (DESIGN.md section 4 "Multi-block": 62..250 outer
    # iterations, occasionally the reference's "Slow progress!" abort, depending on the start): seeds are tried in order
    tried = []
    for seed in range(3):
        X, obj, data = ManiSDP_multiblock(At, b, c, K, dict(tol=1e-8, line_search=1, tau1=1, verbose=False, seed=seed))
        tried.append((seed, data["status"], obj, data["iters"]))
        if data["status"] == 0:
            break
    assert data["status"] == 0 and max(data["gap"], data["pinf"], data["dinf"]) < 1e-8, tried
    assert abs(obj - gold["obj"]) <= 1e-6 * abs(gold["obj"]), tried


def test_multiblock_index_split_matches_the_authors_embedding():
    """data/SDP_demo_1.mat (89 blocks of order 55 / 10 / 10 ...): the (row, column) pairs the engine derives from the
    stacked-vec row indices of At equal the block-diagonal embedding, which tests/test_host_generators.py pins against
    the authors' own single-block form of the same SDP (sedumi.At_full) -- a golden vector of the reference for the
    index handling of the multi-block path (BASELINE north_star: bit-exact A(YY') indices)."""
    import os
    import scipy.sparse as sp
    from instances import generators as G
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdp_demo_1.npz"))
    At = sp.csc_matrix((d["At_data"], d["At_indices"], d["At_indptr"]), shape=tuple(d["At_shape"]))
    At.sort_indices()
    c = np.zeros(At.shape[0])
    c[d["c_idx"]] = d["c_val"]
    ns = [int(v) for v in d["ns"]]
    N = int(sum(ns))
    # entry e of At (CSC order) -> its row in the embedded N*N x m matrix: embed a copy whose values are the entry numbers
    emb = sp.coo_matrix(G.embed_blocks(sp.csc_matrix((np.arange(1, At.nnz + 1, dtype=np.float64), At.indices, At.indptr),
                                                    shape=At.shape), c, {"s": ns}, d["b"])[0])
    order = np.argsort(emb.data)
    rows = emb.row[order]
    with _handle(At, d["b"], c, {"s": ns, "nob": 0}) as h:
        i, j = h.index_split()
    assert len(i) == At.nnz
    assert np.array_equal(i, rows % N) and np.array_equal(j, rows // N)


@pytest.mark.parametrize("p", [[64, 62], [66, 66], [3, 5], [20, 17], [1, 1], [40, 33]])
def test_multiblock_closures_on_long_rows(p):
    """dense blocks give the row-list kernel LONG rows (here ~130 entries per row): the one-CTA-per-row forms of K3
    (affine.cu: k_rowlist_apply_wide for >= 32 vectors per row, k_rowlist_apply_wide_narrow below that) against the
    oracle's dense cell-array closures, and the KKT step (S * identity through the same kernels) against NumPy"""
    from instances import generators as G
    from oracle.manisdp_ref import MultiblockProblem
    from oracle.manopt_rtr import Cells
    At, b, c, K, n, I, coe = G.bqp_sparse_instance(2, 10, 5)
    ns, nob = K["s"], K["nob"]
    assert At.nnz >= 64 * sum(ns)
    rng = np.random.default_rng(23)
    Y = _point(ns, nob, p, rng)
    y = 0.2 * rng.standard_normal(At.shape[1])
    sigma = 1.3
    ora = MultiblockProblem(At.tocsc(), b, c, ns, p, nob, y, sigma)
    Yc = Cells(Y)
    f0 = ora.cost(Yc)
    g0 = ora.grad(Yc)
    U = ora.M.proj(Yc, Cells([rng.standard_normal(B.shape) for B in Y]))
    H0 = ora.hess(Yc, U)
    # dual slack blocks after the dual update, dense on the host
    X = [B @ B.T for B in Y]
    x = np.concatenate([Xi.reshape(-1, order="F") for Xi in X])
    y1 = y - sigma * (At.T @ x - b)
    cy = c - At @ y1
    off2 = np.concatenate([[0], np.cumsum([v * v for v in ns])])
    with _handle(At, b, c, K) as h:
        h.set_dual(y, sigma)
        h.mb_set_Y(Y)
        f = h.cost()
        Gd, gn = h.grad()
        Hd = h.hess(h.mb_join(U.b))
        Gs, Hs = h.mb_split(Gd), h.mb_split(Hd)
        k, dinfs, nneg = h.mb_kkt(update_dual=1)
        for i, nb in enumerate(ns):
            Si = cy[off2[i]:off2[i + 1]].reshape(nb, nb, order="F")
            Si = Si - np.diag(np.sum(X[i] * Si, axis=0))
            vals, vecs = h.mb_block_eigs(i)
            ev = np.linalg.eigvalsh(Si)
            assert np.allclose(vals, ev, rtol=0, atol=1e-11 * max(1.0, np.abs(ev).max()))
            assert np.linalg.norm(Si @ vecs - vecs * vals) < 1e-10 * max(1.0, np.abs(ev).max())
    assert abs(f - f0) <= 1e-12 * max(1.0, abs(f0))
    for i in range(len(ns)):
        assert _rel(Gs[i], g0[i]) < 1e-12
        assert _rel(Hs[i], H0[i]) < 1e-11
