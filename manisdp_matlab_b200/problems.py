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
