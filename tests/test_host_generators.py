"""CPU: instance builders (instances/generators.py) and the multi-block embedding against the reference's own data."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_bqpmom_sizes_match_reference_log():
    """n and m of data/bqp_result.txt:3-8 (q = 10: 56 / 1256, q = 20: 211 / 16361)"""
    from instances import generators as G
    rng = np.random.default_rng(0)
    for q, (n, m) in {10: (56, 1256), 20: (211, 16361)}.items():
        Q = rng.standard_normal((q, q))
        At, b, c, K = G.bqpmom(q, Q + Q.T, rng.standard_normal(q))
        assert (K["s"], At.shape[1]) == (n, m)


def test_embed_blocks_against_reference_single_block_form():
    """multi-block -> single-block index embedding (the map the engine applies at create for MANISDP_MULTIBLOCK handles,
    csrc/affine.cu `split`; its read-back is compared with this embedding on the GPU in
    tests/test_gpu_multiblock.py::test_multiblock_index_split_matches_the_authors_embedding).  data/SDP_demo_1.mat
    carries the authors' own single-block version of the same SDP (sedumi.At_full / c_full): every embedded column must
    appear there, in order, and c must match exactly."""
    import scipy.sparse as sp
    from instances import generators as P
    d = np.load(os.path.join(GOLDEN, "sdp_demo_1.npz"))
    At = sp.csc_matrix((d["At_data"], d["At_indices"], d["At_indptr"]), shape=tuple(d["At_shape"]))
    c = sp.csc_matrix((d["c_val"], (d["c_idx"], np.zeros(len(d["c_idx"]), dtype=int))), shape=(At.shape[0], 1))
    ns = d["ns"]
    Ab, bb, cb, N, off = P.embed_blocks(At, c, {"s": ns}, d["b"])
    assert N == ns.sum() == 2240 and Ab.shape == (N * N, At.shape[1]) and Ab.nnz == At.nnz
    # support is block diagonal and symmetric
    coo = Ab.tocoo()
    i, j = coo.row % N, coo.row // N
    bi = np.searchsorted(off, i, side="right") - 1
    bj = np.searchsorted(off, j, side="right") - 1
    assert np.array_equal(bi, bj)
    ref = "/root/reference/data/SDP_demo_1.mat"
    if os.path.exists(ref):
        import scipy.io as sio
        sed = sio.loadmat(ref, squeeze_me=True, struct_as_record=False)["SDP_1"].sedumi
        Af = sp.csc_matrix(sed.At_full)
        Af.sort_indices()
        Ab.sort_indices()
        assert (sp.csc_matrix(sed.c_full).reshape(-1, 1) != cb).nnz == 0

        def key(M, k):
            s, e = M.indptr[k], M.indptr[k + 1]
            return (M.indices[s:e].tobytes(), np.round(M.data[s:e], 12).tobytes())

        where = {}
        for k in range(Af.shape[1]):
            where.setdefault(key(Af, k), k)
        pos = [where.get(key(Ab, k), -1) for k in range(Ab.shape[1])]
        assert min(pos) >= 0 and all(a < b for a, b in zip(pos, pos[1:]))


def test_bqpsos_sizes_match_the_reference_dual_table():
    """data/bqp_result.txt:20-27 (the authors' ManiDSDP log): d = 10 / 20 / 30 -> n = 56 / 211 / 466 and m = 385 / 6195 /
    31930 equality constraints, i.e. one per non-constant multilinear monomial of degree <= 4 -- the restated bqpsos
    returns that many rows plus the row of the constant term; diag(A*A') counts the positions of every monomial."""
    from instances import generators as G
    rng = np.random.default_rng(0)
    for d, n, m in [(10, 56, 385), (20, 211, 6195), (30, 466, 31930)]:
        Q = rng.standard_normal((d, d))
        A, b, dAAt, mb = G.bqpsos(Q + Q.T, rng.standard_normal(d), d)
        assert (mb, A.shape[0] - 1, A.shape[1]) == (n, m, n * n)
        assert A.nnz == n * n and dAAt.sum() == n * n and dAAt[0] == n  # a partition of the positions
        assert np.array_equal(np.asarray(A.sum(axis=1)).ravel(), dAAt)
