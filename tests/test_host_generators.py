"""CPU: the product-side input builders agree with the oracle's restatement of the reference generators."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_bqpmom_product_equals_oracle():
    from manisdp_matlab_b200 import problems as P
    from oracle import generators as g
    for q in (10, 20):
        d = np.load(os.path.join(GOLDEN, f"bqp_{q}_1.npz"))
        A1, b1, c1, K1 = P.bqpmom(q, d["Q"], d["e"])
        A2, b2, c2, K2 = g.bqpmom(q, d["Q"], d["e"])
        assert K1 == K2 and A1.shape == A2.shape
        assert (A1 != A2).nnz == 0
        assert np.array_equal(np.asarray(b1).ravel(), np.asarray(b2).ravel())
        assert np.array_equal(np.asarray(c1).ravel(), np.asarray(c2).ravel())


def test_bqpmom_sizes_match_reference_log():
    """n and m of data/bqp_result.txt:3-8 (q = 10: 56 / 1256, q = 20: 211 / 16361)"""
    from manisdp_matlab_b200 import problems as P
    rng = np.random.default_rng(0)
    for q, (n, m) in {10: (56, 1256), 20: (211, 16361)}.items():
        Q = rng.standard_normal((q, q))
        At, b, c, K = P.bqpmom(q, Q + Q.T, rng.standard_normal(q))
        assert (K["s"], At.shape[1]) == (n, m)
