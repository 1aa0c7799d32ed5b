"""Sparse-native problem ingest for the MaxCut path: G-set reader, sparse Laplacian (the reference's
src/basicfunction/Laplacian.m:1-12 builds a DENSE L, impossible at n = 1e6) and the synthetic n = 1e6 graphs of
BASELINE.json config 5.  Host-side NumPy/SciPy; nothing here is on the timed path."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def read_gset(path):
    """G-set text file: header `n nedges`, then 1-based `i j w` lines (example/example_maxcut.m:10-17)."""
    with open(path) as fh:
        n, ne = (int(t) for t in fh.readline().split()[:2])
        E = np.loadtxt(fh, ndmin=2)
    E = E[:ne]
    return n, E[:, 0].astype(np.int64) - 1, E[:, 1].astype(np.int64) - 1, E[:, 2].astype(np.float64)


def laplacian(n, ei, ej, w):
    """Sparse graph Laplacian with the semantics of Laplacian.m:5-11: the adjacency entry of a repeated edge is
    ASSIGNED (last one wins, :7) while degrees ACCUMULATE every listed edge (:9-10)."""
    key = np.minimum(ei, ej) * n + np.maximum(ei, ej)
    # last occurrence of each undirected pair
    _, first_rev = np.unique(key[::-1], return_index=True)
    last = len(key) - 1 - first_rev
    a_i, a_j, a_w = ei[last], ej[last], w[last]
    A = sp.coo_matrix((np.concatenate([a_w, a_w]), (np.concatenate([a_i, a_j]), np.concatenate([a_j, a_i]))),
                      shape=(n, n)).tocsr()
    deg = np.bincount(ei, weights=w, minlength=n) + np.bincount(ej, weights=w, minlength=n)
    return (sp.diags(deg) - A).tocsr()


def maxcut_C(n, ei, ej, w):
    """C = -L/4 (example_maxcut.m:18-28)."""
    return (-0.25 * laplacian(n, ei, ej, w)).tocsc()


def synthetic_er(n, mean_degree=48, seed=0):
    """G(n, mean_degree/(n-1))-like graph with unit weights: the 'G1 profile' of SURVEY 8 config 5.  Edges are drawn
    as n*mean_degree/2 uniform pairs (self loops dropped, duplicates merged), which is the sparse limit of G(n,p)."""
    rng = np.random.default_rng(seed)
    ne = int(n * mean_degree // 2)
    ei = rng.integers(0, n, ne, dtype=np.int64)
    ej = rng.integers(0, n, ne, dtype=np.int64)
    keep = ei != ej
    ei, ej = ei[keep], ej[keep]
    key = np.unique(np.minimum(ei, ej) * n + np.maximum(ei, ej))
    ei, ej = key // n, key % n
    return n, ei, ej, np.ones(len(ei))


def synthetic_torus(side, seed=0):
    """side x side toroidal grid with +-1 weights: the G11 / G32 / G81 profile (degree 4)."""
    rng = np.random.default_rng(seed)
    n = side * side
    idx = np.arange(n, dtype=np.int64)
    r, c = idx // side, idx % side
    right = r * side + (c + 1) % side
    down = ((r + 1) % side) * side + c
    ei = np.concatenate([idx, idx])
    ej = np.concatenate([right, down])
    w = rng.choice([-1.0, 1.0], size=len(ei))
    return n, ei, ej, w


def read_sdpa(path):
    """SDPA sparse format (.dat-s, SDPLIB) -> SeDuMi (At, b, c, K) for a single semidefinite block, with the sign
    convention of the reference's reader (src/basicfunction/fromsdpa.m:95,127,141): SDPA's dual
    max <F0,Y> s.t. <Fi,Y> = c_i becomes  min <-F0, X>  s.t. <Fi, X> = c_i,  so  obj = -(SDPLIB optimal value).
    Entries are given for the upper triangle and mirrored; row index of At / c is the column-major j*n + i."""
    with open(path) as fh:
        lines = [ln for ln in fh if ln.strip() and ln.lstrip()[0] not in '"*']
    clean = lambda s: s.translate(str.maketrans("{}(),", "     "))
    m = int(clean(lines[0]).split()[0])
    nblocks = int(clean(lines[1]).split()[0])
    dims = [int(t) for t in clean(lines[2]).split()[:nblocks]]
    if nblocks != 1 or dims[0] <= 0:
        raise ValueError("read_sdpa: only one semidefinite block is supported (K.s scalar, as the four primal drivers)")
    n = dims[0]
    b = np.array([float(t) for t in clean(lines[3]).split()[:m]])
    E = np.array([[float(t) for t in ln.split()[:5]] for ln in lines[4:]])
    mat = E[:, 0].astype(np.int64)
    i = E[:, 2].astype(np.int64) - 1
    j = E[:, 3].astype(np.int64) - 1
    v = E[:, 4]
    off = i != j
    rows = np.concatenate([j * n + i, (i * n + j)[off]])
    cols = np.concatenate([mat, mat[off]])
    vals = np.concatenate([v, v[off]])
    obj = cols == 0
    c = -sp.csc_matrix((vals[obj], (rows[obj], np.zeros(obj.sum(), dtype=np.int64))), shape=(n * n, 1))
    At = sp.csc_matrix((vals[~obj], (rows[~obj], cols[~obj] - 1)), shape=(n * n, m))
    return At, b, c, {"s": n}


# ---------------------------------------------------------------------------------------------------------------------
# SeDuMi-format input of BASELINE config 2 (host-side ingest, never on the timed path): the reference's second-order
# moment relaxation of  min x'Qx + e'x, x in {-1,1}^n  exactly as src/basicfunction/bqpmom.m:6-126 builds it (same
# constraint set, order and weights, so n = 1 + n + n(n-1)/2 and m match data/bqp_result.txt:3-8).  The oracle holds an
# identical restatement; tests/test_host_generators.py checks the two produce the same (At, b, c).
# ---------------------------------------------------------------------------------------------------------------------
import itertools


def _index_map(sp_basis):
    return {tuple(col): k for k, col in enumerate(sp_basis.T.tolist())}


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




# ---------------------------------------------------------------------------------------------------------------------
# multi-block SDPs through the single-block engine
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
