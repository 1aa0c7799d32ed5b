"""INSTANCE BUILDERS (test / benchmark infrastructure, not product code) -- problem generators.

CPU restatement (NumPy/SciPy, FP64, 0-based indices) of the reference's SeDuMi-format problem
builders.  They only produce INPUTS (At, b, c, K / graphs); tests/, tools/, the oracle and bench.py
import them, the product package manisdp_matlab_b200/ does not.

Each function cites the reference file:line it follows (paths relative to the reference root).
All `At` outputs are scipy CSC matrices of shape (n*n, m) whose row index is the column-major
linear index r = j*n + i of X[i, j] (MATLAB `vec`, SURVEY.md section 8c pitfalls).
"""
from __future__ import annotations

import itertools
from math import comb

import numpy as np
import scipy.sparse as sp


# --------------------------------------------------------------------------------------------
# monomial bases
# --------------------------------------------------------------------------------------------
def get_basis_sequential(n: int, d: int) -> np.ndarray:
    """Literal restatement of src/basicfunction/get_basis.m:1-33 (successor rule, one column at a
    time).  Slow; used only to pin `get_basis` (the sorted construction) in the CPU tests."""
    lb = comb(n + d, d)
    basis = np.zeros((n, lb), dtype=np.int64)
    i = 0
    t = 0  # 0-based column of the last written monomial
    while i < d + 1:
        t += 1
        if basis[n - 1, t - 1] == i:
            if i < d:
                basis[0, t] = i + 1
            i += 1
        else:
            j = 0
            while basis[j, t - 1] == 0:
                j += 1
            basis[:, t] = basis[:, t - 1]
            if j == 0:
                basis[0, t] -= 1
                basis[1, t] += 1
            else:
                basis[0, t] = basis[j, t] - 1
                basis[j, t] = 0
                basis[j + 1, t] += 1
    return basis


def get_basis(n: int, d: int) -> np.ndarray:
    """Exponent vectors of all monomials of degree <= d in n variables, as columns, in the order
    produced by src/basicfunction/get_basis.m:1-33: graded by total degree, ties broken by the
    exponent of x_n, then x_{n-1}, ... (the order `comp.m:1-24` defines, which `bfind.m` relies on).
    Built by enumeration + sort; `get_basis_sequential` pins it."""
    cols = []
    for deg in range(d + 1):
        for combo in itertools.combinations_with_replacement(range(n), deg):
            e = np.zeros(n, dtype=np.int64)
            for v in combo:
                e[v] += 1
            cols.append(e)
    B = np.array(cols, dtype=np.int64)  # (lb, n)
    # sort key: (degree, e[n-1], e[n-2], ..., e[0]); np.lexsort uses the LAST key as primary
    keys = [B[:, v] for v in range(n)] + [B.sum(axis=1)]
    order = np.lexsort(keys)
    return np.ascontiguousarray(B[order].T)


def _index_map(sp_basis: np.ndarray) -> dict:
    """Dictionary replacement for the binary search src/basicfunction/bfind.m:1-20."""
    return {tuple(col): k for k, col in enumerate(sp_basis.T.tolist())}


# --------------------------------------------------------------------------------------------
# MaxCut
# --------------------------------------------------------------------------------------------
def read_gset(path: str):
    """Parse a G-set text file: header `nv ne`, then `i j w` lines (1-based)."""
    with open(path) as fh:
        toks = fh.read().split()
    nv, ne = int(toks[0]), int(toks[1])
    arr = np.array(toks[2 : 2 + 3 * ne], dtype=np.float64).reshape(ne, 3)
    return nv, arr[:, 0].astype(np.int64) - 1, arr[:, 1].astype(np.int64) - 1, arr[:, 2].copy()


def laplacian(nv: int, ei: np.ndarray, ej: np.ndarray, w: np.ndarray) -> sp.csr_matrix:
    """Graph Laplacian with the semantics of src/basicfunction/Laplacian.m:1-12: off-diagonal
    entries are ASSIGNED (a duplicate edge overwrites: last wins, line 7-8) while the degrees
    ACCUMULATE over every listed edge (lines 9-10).  Returned sparse (the reference builds a dense
    L and calls sparse() on it, example/example_maxcut.m:10-11)."""
    off = {}
    deg = np.zeros(nv)
    for a, b, ww in zip(ei.tolist(), ej.tolist(), w.tolist()):
        off[(a, b)] = -ww
        off[(b, a)] = -ww
        deg[a] += ww
        deg[b] += ww
    if off:
        keys = np.array(list(off.keys()), dtype=np.int64)
        vals = np.array(list(off.values()))
        rows = np.concatenate([keys[:, 0], np.arange(nv)])
        cols = np.concatenate([keys[:, 1], np.arange(nv)])
        data = np.concatenate([vals, deg])
    else:
        rows = cols = np.arange(nv)
        data = deg
    # a self loop (a == a) would have been assigned then incremented in the reference; G-set has none
    L = sp.coo_matrix((data, (rows, cols)), shape=(nv, nv)).tocsr()
    L.sum_duplicates()
    L.eliminate_zeros()  # sparse(L) drops explicit zeros
    return L


def maxcut_C(nv, ei, ej, w) -> sp.csr_matrix:
    """C = -L/4 (example/example_maxcut.m:10-11)."""
    return (-0.25 * laplacian(nv, ei, ej, w)).tocsr()


def maxcut_sedumi(C: sp.spmatrix):
    """The (At, b, c, K) description of diag(X)=1 used in example/example_maxcut.m:12-21."""
    n = C.shape[0]
    rows = np.arange(n) * n + np.arange(n)
    At = sp.csc_matrix((np.ones(n), (rows, np.arange(n))), shape=(n * n, n))
    c = np.asarray(C.todense()).reshape(-1, order="F")
    return At, np.ones(n), c, {"s": n}


# --------------------------------------------------------------------------------------------
# BQP second-order moment relaxation
# --------------------------------------------------------------------------------------------
def bqpmom(n: int, Q: np.ndarray, e: np.ndarray):
    """Restates src/basicfunction/bqpmom.m:6-126.  Returns (At, b, c, K) with At CSC (mb^2 x m),
    b dense (m,), c dense (mb^2,), K = {'s': mb}."""
    basis = get_basis(n, 2)
    basis = basis[:, (basis > 1).sum(axis=0) == 0]  # bqpmom.m:8-14  multilinear monomials
    mb = basis.shape[1]
    spb = get_basis(n, 4)
    keep = ((spb > 2).sum(axis=0) == 0) & ((spb % 2).sum(axis=0) != 0)  # bqpmom.m:16-22
    spb = spb[:, keep]
    lsp = spb.shape[1]
    where = _index_map(spb)
    mm = [[] for _ in range(lsp)]  # mm[ind] = list of (i, j), i < j, 0-based   bqpmom.m:24-31
    bt = basis.T
    for i in range(mb):
        s = bt[i] + bt[i + 1 :]
        for off, col in enumerate(s.tolist()):
            mm[where[tuple(col)]].append((i, i + 1 + off))
    ncons = mb * (mb + 1) // 2 - lsp + n * (mb - 1) - mb + 1  # bqpmom.m:32
    row, col, val = [0], [0], [1.0]  # X(1,1) = 1, bqpmom.m:33-37
    b = np.zeros(ncons)
    b[0] = 1.0
    for i in range(1, n + 1):  # bqpmom.m:38-42
        row += [0, i * mb + i]
        col += [i, i]
        val += [0.5, -0.5]
    l = n + 1
    for i in range(n + 1, mb):  # bqpmom.m:45-51
        cc = np.nonzero(basis[:, i] == 1)[0] + 1
        row += [cc[0] * mb + cc[0], i * mb + i, cc[1] * mb + cc[1], i * mb + i]
        col += [l, l, l + 1, l + 1]
        val += [0.5, -0.5, 0.5, -0.5]
        l += 2
    loa = []  # bqpmom.m:52-58 : both symmetric positions of every pair
    for i in range(lsp):
        a = []
        for (p, q) in mm[i]:
            a += [q * mb + p, p * mb + q]
        loa.append(a)
    for k in range(n):  # bqpmom.m:59-78   x_k^2 * m_i = m_i
        for i in range(1, mb):
            if basis[k, i] == 0:
                bi = basis[:, i].copy()
                bi[k] = 2
                l1 = loa[where[tuple(bi.tolist())]]
                l2 = loa[where[tuple(basis[:, i].tolist())]]
                row += l1 + l2
                col += [l] * (len(l1) + len(l2))
                if len(l1) < len(l2):
                    val += [1.0] * len(l1) + [-len(l1) / len(l2)] * len(l2)
                else:
                    val += [len(l2) / len(l1)] * len(l1) + [-1.0] * len(l2)
                l += 1
    for i in range(lsp):  # bqpmom.m:80-90  entries of one monomial are all equal
        firsts = [p for (p, _) in mm[i]]
        idx = int(np.argmax(firsts))
        for j in range(len(mm[i])):
            if j != idx:
                row += loa[i][2 * idx : 2 * idx + 2] + loa[i][2 * j : 2 * j + 2]
                col += [l] * 4
                val += [0.5, 0.5, -0.5, -0.5]
                l += 1
    assert l == ncons, (l, ncons)
    At = sp.coo_matrix((val, (row, col)), shape=(mb * mb, ncons)).tocsc()
    At.sum_duplicates()
    # objective, bqpmom.m:93-122
    crow = list(range(1, n + 1))
    ccol = list(range(1, n + 1))
    cval = list(np.diag(Q))
    for i in range(n):
        cnt = len(mm[i])
        for (p, q) in mm[i]:
            crow += [p, q]
            ccol += [q, p]
        cval += [e[i] / (2 * cnt)] * (2 * cnt)
    ind = n
    for i in range(1, n):
        for j in range(i):
            cnt = len(mm[ind])
            for (p, q) in mm[ind]:
                crow += [p, q]
                ccol += [q, p]
            cval += [Q[j, i] / cnt] * (2 * cnt)
            ind += 1
    C = sp.coo_matrix((cval, (crow, ccol)), shape=(mb, mb)).toarray()
    c = C.reshape(-1, order="F")
    return At, b, c, {"s": mb}


def bqp_bruteforce(Q: np.ndarray, e: np.ndarray) -> float:
    """min x'Qx + e'x over x in {-1,+1}^n (independent check for small n)."""
    n = len(e)
    X = np.array(list(itertools.product([-1.0, 1.0], repeat=n)))
    vals = np.einsum("bi,ij,bj->b", X, Q, X) + X @ e
    return float(vals.min())


# --------------------------------------------------------------------------------------------
# quartic on the sphere, second-order moment relaxation
# --------------------------------------------------------------------------------------------
def qsmom(n: int, coe: np.ndarray):
    """Restates src/basicfunction/qsmom.m:6-124."""
    basis = get_basis(n, 2)
    mb = basis.shape[1]
    spb = get_basis(n, 4)
    lsp = spb.shape[1]
    assert len(coe) == lsp
    where = _index_map(spb)
    mm = [[] for _ in range(lsp)]  # qsmom.m:11-18, i <= j
    bt = basis.T
    for i in range(mb):
        s = bt[i] + bt[i:]
        for off, colv in enumerate(s.tolist()):
            mm[where[tuple(colv)]].append((i, i + off))
    ncons = mb * (mb + 1) // 2 - lsp + mb + 1  # qsmom.m:19
    row, col, val = [0], [0], [1.0]
    b = np.zeros(ncons)
    b[0] = 1.0
    l = 1

    def entries(ind):
        """positions contributed by mm[ind]: diagonal pairs once, off-diagonal pairs twice
        (qsmom.m:39-47).  Order: loa(2j-1:2j) = [(q,p)->q*mb+p ... ] as in qsmom.m:28-31."""
        out = []
        for (p, q) in mm[ind]:
            if p == q:
                out.append(p * mb + q)
            else:
                out += [q * mb + p, p * mb + q]
        return out

    ent_cache = [entries(i) for i in range(lsp)]
    eye = np.eye(n, dtype=np.int64)
    for i in range(mb):  # qsmom.m:33-64   (sum_k x_k^2) * m_i = m_i
        for k in range(n):
            ind1 = where[tuple((basis[:, i] + 2 * eye[k]).tolist())]
            e1 = ent_cache[ind1]
            row += e1
            col += [l] * len(e1)
            val += [1.0 / len(e1)] * len(e1)
        e2 = ent_cache[where[tuple(basis[:, i].tolist())]]
        row += e2
        col += [l] * len(e2)
        val += [-1.0 / len(e2)] * len(e2)
        l += 1
    for i in range(lsp):  # qsmom.m:66-92
        firsts = [p for (p, _) in mm[i]]
        idx = int(np.argmax(firsts))
        pi, qi = mm[i][idx]
        for j in range(len(mm[i])):
            if j == idx:
                continue
            if pi == qi:
                row.append(pi * mb + qi); col.append(l); val.append(1.0)
            else:
                row += [qi * mb + pi, pi * mb + qi]; col += [l, l]; val += [0.5, 0.5]
            pj, qj = mm[i][j]
            if pj == qj:
                row.append(pj * mb + qj); col.append(l); val.append(-1.0)
            else:
                row += [qj * mb + pj, pj * mb + qj]; col += [l, l]; val += [-0.5, -0.5]
            l += 1
    assert l == ncons, (l, ncons)
    At = sp.coo_matrix((val, (row, col)), shape=(mb * mb, ncons)).tocsc()
    At.sum_duplicates()
    crow, ccol, cval = [], [], []  # qsmom.m:95-112
    for i in range(lsp):
        s = 0
        for (p, q) in mm[i]:
            if p == q:
                crow.append(p); ccol.append(q); s += 1
            else:
                crow += [p, q]; ccol += [q, p]; s += 2
        cval += [coe[i] / s] * s
    C = sp.coo_matrix((cval, (crow, ccol)), shape=(mb, mb)).toarray()
    return At, b, C.reshape(-1, order="F"), {"s": mb}


# --------------------------------------------------------------------------------------------
# Lovasz theta
# --------------------------------------------------------------------------------------------
def generate_hamming(k: int, d):
    """Restates example/generate_hamming.m:25-60.  Returns (At, b, c, K); constraint 0 is the trace
    row (generate_hamming.m:55), the others one per edge in (vertex, bit pattern) order."""
    n = 1 << k
    d = [d] if np.isscalar(d) else list(d)
    bitpat = []
    for dist in d:  # generate_hamming.m:31-37 : nchoosek(1:k,i) is lexicographic
        for combo in itertools.combinations(range(k), dist):
            bitpat.append(sum(1 << bpos for bpos in combo))
    bitpat = np.array(bitpat, dtype=np.int64)
    ai, aj = [], []
    adj_r, adj_c = [], []
    start = 1  # row 0 = trace
    for i in range(n):
        nb = np.bitwise_xor(i, bitpat)
        nb = nb[nb > i]
        if len(nb):
            adj_r += [i] * len(nb) + nb.tolist()
            adj_c += nb.tolist() + [i] * len(nb)
            rows = np.arange(start, start + len(nb))
            ai += rows.tolist() + rows.tolist()
            aj += (nb * n + i).tolist() + (nb + i * n).tolist()
            start += len(nb)
    m = start
    ai = np.array(ai + [0] * n, dtype=np.int64)
    aj = np.array(aj + (np.arange(n) * n + np.arange(n)).tolist(), dtype=np.int64)
    At = sp.coo_matrix((np.ones(len(ai)), (aj, ai)), shape=(n * n, m)).tocsc()
    Adj = sp.coo_matrix((np.ones(len(adj_r)), (adj_r, adj_c)), shape=(n, n)).toarray()
    Adj = (Adj > 0).astype(np.float64)
    c = -(1.0 - Adj).reshape(-1, order="F")  # generate_hamming.m:56
    b = np.zeros(m)
    b[0] = 1.0
    return At, b, c, {"s": n}


def theta_random(n: int, nedges_draw: int, seed: int):
    """example/example_theta.m:2-44 with NumPy's generator in place of MATLAB's rng(1)/randi
    (MATLAB streams are not reproducible here): C = -J, one constraint X_ij = 0 per sampled edge
    (i < j, unique) and the trace row LAST (example_theta.m:36-43)."""
    rng = np.random.default_rng(seed)
    om = rng.integers(0, n, size=(nedges_draw, 2))
    om = om[om[:, 0] < om[:, 1]]
    om = np.unique(om, axis=0)
    m = len(om)
    rows = np.concatenate([om[:, 0] * n + om[:, 1], om[:, 1] * n + om[:, 0], np.arange(n) * n + np.arange(n)])
    cols = np.concatenate([np.arange(m), np.arange(m), np.full(n, m)])
    At = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(n * n, m + 1)).tocsc()
    b = np.zeros(m + 1)
    b[m] = 1.0
    c = -np.ones(n * n)
    return At, b, c, {"s": n}


# --------------------------------------------------------------------------------------------
# multi-block instances (inputs of ManiSDP_multiblock: K.s = block orders, K.nob = number of
# leading unit-diagonal blocks; At has sum(n_i^2) rows, the stacked column-major vecs of the blocks)
# --------------------------------------------------------------------------------------------
def multiblock_random(nset, nob: int, m: int, seed: int, rank: int = 2):
    """Random feasible, bounded multi-block SDP in the format ManiSDP_multiblock.m:7 takes.
    Feasible point: X_i = V_i V_i' with `rank` columns (unit rows for the first nob blocks);
    constraints: m sparse symmetric matrices, each supported on one or two blocks, b = A(X0);
    objective: a positive definite matrix per block, so the problem is bounded below."""
    rng = np.random.default_rng(seed)
    nset = [int(v) for v in nset]
    off = np.concatenate([[0], np.cumsum([v * v for v in nset])]).astype(np.int64)
    x0 = []
    for i, n in enumerate(nset):
        V = rng.standard_normal((n, min(rank, n)))
        if i < nob:
            V /= np.linalg.norm(V, axis=1, keepdims=True)
        x0.append((V @ V.T).reshape(-1, order="F"))
    x0 = np.concatenate(x0)
    rows, cols, vals = [], [], []
    for k in range(m):
        for blk in rng.choice(len(nset), size=min(len(nset), int(rng.integers(1, 3))), replace=False):
            n = nset[blk]
            for _ in range(int(rng.integers(1, 4))):
                a, bb = int(rng.integers(0, n)), int(rng.integers(0, n))
                v = float(rng.standard_normal())
                if a == bb:
                    if blk < nob:
                        continue  # the diagonal of a unit-diagonal block is fixed by the manifold
                    rows.append(off[blk] + a * n + a); cols.append(k); vals.append(v)
                else:
                    rows += [off[blk] + bb * n + a, off[blk] + a * n + bb]
                    cols += [k, k]
                    vals += [0.5 * v, 0.5 * v]
    At = sp.csc_matrix((vals, (rows, cols)), shape=(int(off[-1]), m))
    At.sum_duplicates()
    b = At.T @ x0
    c = []
    for n in nset:
        G = rng.standard_normal((n, n))
        c.append((G @ G.T / n + 0.5 * np.eye(n)).reshape(-1, order="F"))
    return At, b, np.concatenate(c), {"s": nset, "nob": int(nob)}


def embed_multiblock(At, c, K):
    """Block-diagonal embedding of a multi-block problem into ONE PSD cone of order N = sum(n_i):
    returns (At_big (N*N, m), c_big (N*N,), N, row offsets).  The off-diagonal blocks of X carry no
    cost and no constraint, and every principal block of a PSD matrix is PSD, so the optimum over
    {X >= 0 of order N} equals the multi-block optimum when no block is unit-diagonal (K.nob = 0) --
    an independent route to the optimum through the single-block drivers (test pin)."""
    nset = [int(v) for v in np.atleast_1d(K["s"])]
    N = int(sum(nset))
    off2 = np.concatenate([[0], np.cumsum([v * v for v in nset])]).astype(np.int64)
    roff = np.concatenate([[0], np.cumsum(nset)]).astype(np.int64)
    mp = np.empty(int(off2[-1]), dtype=np.int64)
    for i, n in enumerate(nset):
        loc = np.arange(n * n, dtype=np.int64)
        a, bb = loc % n, loc // n
        mp[off2[i]:off2[i + 1]] = (roff[i] + bb) * N + (roff[i] + a)
    At = sp.coo_matrix(At)
    At_big = sp.csc_matrix((At.data, (mp[At.row], At.col)), shape=(N * N, At.shape[1]))
    c = np.asarray(c, dtype=np.float64).ravel()
    c_big = np.zeros(N * N)
    c_big[mp] = c
    return At_big, c_big, N, roff


def get_basis_on(n: int, d: int, var) -> np.ndarray:
    """get_basis(n, d, var) of src/basicfunction/get_basis.m:1-33 with the optional variable list: the monomials of
    degree <= d in the variables `var` (0-based, ascending), as exponent vectors over all n variables, in the order the
    successor rule produces (the order of `get_basis` on len(var) variables)."""
    var = [int(v) for v in var]
    sub = get_basis(len(var), d)
    out = np.zeros((n, sub.shape[1]), dtype=np.int64)
    out[var, :] = sub
    return out


def bqpmom_sparse(n: int, I, coe: np.ndarray):
    """Restates src/basicfunction/bqpmom_sparse.m:6-132: second-order moment relaxation of a BQP whose objective is a
    sum of quadratics on the cliques I[0..t) (0-based variable lists, ascending), one PSD block per clique.
    coe: coefficients of the non-constant multilinear monomials of degree <= 2 supported on a clique, in the
    lexicographic (sortrows) order of their exponent vectors (example/example_bqp_sparse.m:9-17).
    Returns (At, b, c, K) with At CSC (sum(mb^2) x ncons), K = {'s': [mb_1..mb_t]}; the caller sets K['nob']."""
    t = len(I)
    basis, mb, spl = [], [], []
    for k in range(t):  # :10-28
        bk = get_basis_on(n, 2, I[k])
        bk = bk[:, (bk > 1).sum(axis=0) == 0]
        basis.append(bk)
        mb.append(bk.shape[1])
        tmp = get_basis_on(n, 4, I[k])
        keep = ((tmp > 2).sum(axis=0) == 0) & ((tmp % 2).sum(axis=0) != 0)
        spl.append(tmp[:, keep])
    spb = np.unique(np.concatenate(spl, axis=1).T, axis=0).T  # :29-30 unique + sortrows: lexicographic in x_1, x_2, ...
    lsp = spb.shape[1]
    where = _index_map(spb)
    off = np.concatenate([[0], np.cumsum([v * v for v in mb])]).astype(np.int64)
    mm = [[] for _ in range(lsp)]  # (i, j, k), i < j inside block k   :33-41
    for k in range(t):
        bt = basis[k].T
        for i in range(mb[k]):
            s = bt[i] + bt[i + 1:]
            for o_, col in enumerate(s.tolist()):
                mm[where[tuple(col)]].append((i, i + 1 + o_, k))
    mc = [len(v) for v in I]
    ncons = sum(v * (v + 1) // 2 for v in mb) - lsp + sum(a * (v - 1) for a, v in zip(mc, mb)) - sum(mb) + t  # :46
    row, col, val = [0], [0], [1.0]  # :47-51
    l = 1
    for k in range(t):  # :52-65  diagonal of the degree-<=1 part equals X_1(1,1)
        for i in range(1 if k == 0 else 0, mc[k] + 1):
            row += [0, off[k] + i * mb[k] + i]
            col += [l, l]
            val += [0.5, -0.5]
            l += 1
    for k in range(t):  # :66-76  X_k(ij,ij) = X_k(i,i) = X_k(j,j)
        pos = {v: a for a, v in enumerate(I[k])}
        for i in range(mc[k] + 1, mb[k]):
            v2 = np.nonzero(basis[k][:, i] == 1)[0]
            c1, c2 = pos[int(v2[0])] + 1, pos[int(v2[1])] + 1
            row += [off[k] + c1 * mb[k] + c1, off[k] + i * mb[k] + i, off[k] + c2 * mb[k] + c2, off[k] + i * mb[k] + i]
            col += [l, l, l + 1, l + 1]
            val += [0.5, -0.5, 0.5, -0.5]
            l += 2
    loa = []  # :77-84
    for i in range(lsp):
        a = []
        for (p_, q_, k) in mm[i]:
            a += [off[k] + q_ * mb[k] + p_, off[k] + p_ * mb[k] + q_]
        loa.append(a)
    for q in range(t):  # :85-104   x_k^2 * m_i = m_i
        for k in range(mc[q]):
            v = I[q][k]
            for i in range(1, mb[q]):
                if basis[q][v, i] == 0:
                    bi = basis[q][:, i].copy()
                    bi[v] = 2
                    l1 = loa[where[tuple(bi.tolist())]]
                    l2 = loa[where[tuple(basis[q][:, i].tolist())]]
                    row += l1 + l2
                    col += [l] * (len(l1) + len(l2))
                    if len(l1) < len(l2):
                        val += [1.0] * len(l1) + [-len(l1) / len(l2)] * len(l2)
                    else:
                        val += [len(l2) / len(l1)] * len(l1) + [-1.0] * len(l2)
                    l += 1
    for i in range(lsp):  # :106-116
        idx = int(np.argmax([p_ for (p_, _, _) in mm[i]]))
        for j in range(len(mm[i])):
            if j != idx:
                row += loa[i][2 * idx:2 * idx + 2] + loa[i][2 * j:2 * j + 2]
                col += [l] * 4
                val += [0.5, 0.5, -0.5, -0.5]
                l += 1
    assert l == ncons, (l, ncons)
    At = sp.coo_matrix((val, (np.asarray(row, dtype=np.int64), col)), shape=(int(off[-1]), ncons)).tocsc()
    At.sum_duplicates()
    b = np.zeros(ncons)
    b[0] = 1.0
    nsp = spb[:, spb.sum(axis=0) <= 2]  # :119-125
    nsp = nsp[:, (nsp > 1).sum(axis=0) == 0]
    assert nsp.shape[1] == len(coe), (nsp.shape, len(coe))
    c = np.zeros(int(off[-1]))
    for i in range(nsp.shape[1]):  # :126-130
        a = loa[where[tuple(nsp[:, i].tolist())]]
        c[a] = coe[i] / len(a)
    return At, b, c, {"s": [int(v) for v in mb]}


def bqp_sparse_instance(t: int, q: int, seed: int):
    """The clique structure of example/example_bqp_sparse.m:4-17 (t cliques of q variables, consecutive cliques share two)
    with N(0,1) coefficients from NumPy's generator.  Returns (At, b, c, K, n, I, coe) with K['nob'] = t."""
    n = q + (q - 2) * (t - 1)
    I = [list(range((q - 2) * i, (q - 2) * (i + 1) + 2)) for i in range(t)]
    mono = set()
    for Ik in I:
        bk = get_basis_on(n, 2, Ik)
        bk = bk[:, (bk > 1).sum(axis=0) == 0]
        mono.update(tuple(cl) for cl in bk.T.tolist())
    coe = np.random.default_rng(seed).standard_normal(len(mono) - 1)
    At, b, c, K = bqpmom_sparse(n, I, coe)
    K["nob"] = t
    return At, b, c, K, n, I, coe


def bqp_sparse_bruteforce(n: int, I, coe: np.ndarray) -> float:
    """min over x in {-1,+1}^n of the clique-sparse quadratic whose coefficients `coe` follow the monomial order of
    bqpmom_sparse (lexicographic exponent vectors of the non-constant multilinear monomials of degree <= 2)."""
    mono = set()
    for Ik in I:
        bk = get_basis_on(n, 2, Ik)
        bk = bk[:, (bk > 1).sum(axis=0) == 0]
        mono.update(tuple(cl) for cl in bk.T.tolist())
    mono.discard(tuple([0] * n))
    mono = sorted(mono)
    E = np.array(mono, dtype=np.int64)  # (nmono, n)
    best = np.inf
    for bits in range(1 << n):
        x = np.array([1.0 if (bits >> v) & 1 else -1.0 for v in range(n)])
        vals = np.prod(np.where(E == 1, x[None, :], 1.0), axis=1)
        best = min(best, float(coe @ vals))
    return best


# --------------------------------------------------------------------------------------------
# SOS (dual) form of the BQP relaxation, input of ManiDSDP_unitdiag
# --------------------------------------------------------------------------------------------
def bqpsos(Q: np.ndarray, e: np.ndarray, n: int):
    """Restates src/basicfunction/bqpsos.m:7-39: second-order SOS relaxation of min x'Qx + e'x, x_i^2 = 1.
    Returns (A, b, dAAt, mb): A CSR (lsp x mb^2) whose rows partition the positions of the mb x mb Gram matrix by the
    multilinear monomial they represent (x_i^2 = 1 reduces exponents mod 2), b the coefficient of each monomial,
    dAAt = diag(A*A') = the number of positions of each monomial."""
    spb = get_basis(n, 4)
    spb = spb[:, (spb > 1).sum(axis=0) == 0]  # :8-11 multilinear monomials of degree <= 4
    mb = comb(n + 2, 2) - n  # :12
    lsp = spb.shape[1]
    where = _index_map(spb)
    rows = np.zeros(mb * mb, dtype=np.int64)
    cols = np.zeros(mb * mb, dtype=np.int64)
    dAAt = np.zeros(lsp)
    dAAt[0] = mb
    cols[:mb] = np.arange(mb) * mb + np.arange(mb)  # :18 the diagonal represents the constant monomial
    ind = mb
    bt = spb[:, :mb].T
    for i in range(mb):  # :20-31
        s = (bt[i] + bt[i + 1:]) % 2
        for o_, colv in enumerate(s.tolist()):
            j = i + 1 + o_
            locb = where[tuple(colv)]
            rows[ind] = rows[ind + 1] = locb
            cols[ind] = i * mb + j
            cols[ind + 1] = j * mb + i
            dAAt[locb] += 2
            ind += 2
    A = sp.csr_matrix((np.ones(mb * mb), (rows, cols)), shape=(lsp, mb * mb))
    b = np.zeros(lsp)  # :34-37
    b[0] = np.trace(Q)
    b[1:n + 1] = e
    iu = np.triu(np.ones((n, n)), 1) != 0
    b[n + 1:(n + 1) * (n + 2) // 2 - n] = 2 * Q.T[iu.T]  # MATLAB's column-major logical indexing of triu(.,1)
    return A, b, dAAt, mb


# ---------------------------------------------------------------------------------------------------------------------
# multi-block SDPs embedded into one block (test infrastructure: pins the block-aware index split of the engine)
# ---------------------------------------------------------------------------------------------------------------------
def embed_blocks(At, c, K, b=None, nob=0):
    """SeDuMi data of a multi-block SDP (K['s'] = [n_1, ..., n_t], rows of At / c index the concatenation of the
    column-major vec(X_i)) -> data of ONE block of order N = sum n_i whose cost and constraints only touch the diagonal
    blocks.  The two SDPs have the same optimum and the same dual slack (block diag(S_i)): off-diagonal blocks of X are
    neither priced nor constrained, and X >= 0 iff it can be completed from PSD diagonal blocks.  The first `nob`
    blocks are unit-diagonal in the reference driver (src/primal/ManiSDP_multiblock.m:1-5); here those diagonal
    constraints are appended to (At, b) explicitly.  Returns (At_big, b_big, c_big, N, offsets)."""
    ns = np.atleast_1d(np.asarray(K["s"] if isinstance(K, dict) else K)).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(ns)])            # row/column offset of block i in the big matrix
    voff = np.concatenate([[0], np.cumsum(ns * ns)])      # offset of vec(X_i) in the concatenated vector
    N = int(off[-1])

    def remap(r):
        r = np.asarray(r, dtype=np.int64)
        blk = np.searchsorted(voff, r, side="right") - 1
        loc = r - voff[blk]
        i = loc % ns[blk]
        j = loc // ns[blk]
        return (off[blk] + j) * N + (off[blk] + i)

    At = sp.coo_matrix(At)
    At_big = sp.csc_matrix((At.data, (remap(At.row), At.col)), shape=(N * N, At.shape[1]))
    cc = sp.coo_matrix(c.reshape(-1, 1) if not sp.issparse(c) else sp.csc_matrix(c).reshape(-1, 1))
    c_big = sp.csc_matrix((cc.data, (remap(cc.row), np.zeros(cc.nnz, dtype=np.int64))), shape=(N * N, 1))
    b_big = None if b is None else (np.asarray(b.todense()).ravel() if sp.issparse(b) else np.asarray(b, float).ravel())
    if nob > 0:
        d = np.arange(off[nob], dtype=np.int64)
        D = sp.csc_matrix((np.ones(len(d)), (d * N + d, np.arange(len(d)))), shape=(N * N, len(d)))
        At_big = sp.hstack([At_big, D], format="csc")
        if b_big is not None:
            b_big = np.concatenate([b_big, np.ones(len(d))])
    return At_big, b_big, c_big, N, off
