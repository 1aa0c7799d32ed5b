"""Pins the optima of the BASELINE configurations at their STATED sizes with the oracle (CPU, minutes each):

  * config 3: quartic on the sphere, q = 60 (n = 1891, m = 1 155 402).  data/qs_c_60_*.txt is absent from the reference
    tree (.MISSING_LARGE_BLOBS), so the 635 376 coefficients are N(0,1) draws of numpy's default_rng(60); their SHA-256
    is recorded so that the GPU test proves it solved the same instance.  Options of example/example_qsphere.m:21-27.
  * config 2: BQP q = 60 (n = 1831, m = 1 155 281), data/bqp_{Q,e}_60_1.txt, options of example/example_bqp.m:31-41.

    python tests/golden/make_golden_large.py [qs60] [bqp60] [bqpsparse] [bqpdual60]

  * multi-block: the sparse BQP of example/example_bqp_sparse.m (20 blocks of order 211) through ManiSDP_multiblock.

Results are merged into tests/golden/oracle_outputs_large.json.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from instances import generators as g  # noqa: E402
from oracle import manisdp_ref as ref  # noqa: E402

OUT = f"{HERE}/oracle_outputs_large.json"


def qs60_coefficients():
    coe = np.random.default_rng(60).standard_normal(635376)
    return coe, hashlib.sha256(coe.tobytes()).hexdigest()


def merge(key, val):
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    out[key] = val
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1)
    print(key, json.dumps(val), flush=True)


def qs60():
    coe, sha = qs60_coefficients()
    At, b, c, K = g.qsmom(60, coe)
    t0 = time.perf_counter()
    opts = dict(tol=1e-8, theta=1e-2, tau1=0.02, delta=6)
    X, obj, data = ref.ManiSDP(At, b, c, K, dict(opts, seed=0))
    merge("qs_c_60_rng60_opt", dict(options=opts, status=int(data["status"]), obj=obj, eta=max(data["gap"], data["pinf"], data["dinf"]), hv=data["hv_count"],
                                    iters=data["iters"], n=int(K["s"]), m=int(At.shape[1]), coe_sha256=sha,
                                    oracle_seconds=time.perf_counter() - t0))


def bqp60():
    d = np.load(f"{HERE}/bqp_60_1.npz")
    At, b, c, K = g.bqpmom(60, d["Q"], d["e"])
    mc = float(np.abs(c).max())
    t0 = time.perf_counter()
    X, obj, data = ref.ManiSDP_unitdiag(At, b, c / mc, K, dict(seed=0))
    merge("bqp_60_1_opt", dict(obj_scaled=obj, obj=obj * mc, eta=max(data["gap"], data["pinf"], data["dinf"]),
                               hv=data["hv_count"], iters=data["iters"], n=int(K["s"]), m=int(At.shape[1]),
                               oracle_seconds=time.perf_counter() - t0))


def bqpsparse():
    """multi-block: example/example_bqp_sparse.m:4-29 at its stated size (t = 20 cliques of q = 20 variables), N(0,1)
    coefficients of numpy's default_rng(1) in place of MATLAB's rng(1) stream"""
    At, b, c, K, n, I, coe = g.bqp_sparse_instance(20, 20, 1)
    opts = dict(tol=1e-8, line_search=1, tau1=1)
    t0 = time.perf_counter()
    X, obj, data = ref.ManiSDP_multiblock(At, b, c, K, dict(opts, seed=0))
    merge("bqp_sparse_20_20", dict(options=opts, status=int(data["status"]), obj=obj,
                                   eta=max(data["gap"], data["pinf"], data["dinf"]), hv=data["hv_count"],
                                   iters=data["iters"], blocks=len(K["s"]), block_order=int(K["s"][0]), m=int(At.shape[1]),
                                   nnz=int(At.nnz), coe_sha256=hashlib.sha256(coe.tobytes()).hexdigest(),
                                   oracle_seconds=time.perf_counter() - t0))


def bqpdual60():
    """dual approach on config 2's instance: example/dual/example_bqp_dual.m:19-36 on data/bqp_{Q,e}_60_1.txt
    (SOS form, n = 1831, m = 523 686 monomials); the optimum equals the primal KAT -520.38067984 (strong duality)"""
    import scipy.sparse as sp
    d = np.load(f"{HERE}/bqp_60_1.npz")
    A, b, dAAt, mb = g.bqpsos(d["Q"], d["e"], 60)
    v = np.zeros((A.shape[0], 1))
    v[0] = 1.0
    A2 = sp.hstack([sp.csr_matrix(v), A]).tocsr()
    c = np.concatenate([[1.0], np.zeros(mb * mb)])
    maxb = float(np.abs(b).max())
    opts = dict(tol=1e-8, line_search=1)
    t0 = time.perf_counter()
    X, obj, data = ref.ManiDSDP_unitdiag(A2, b / maxb, c, {"f": 1, "s": mb}, dict(opts, dAAt=dAAt, seed=0))
    merge("bqp_60_1_dual", dict(options=opts, status=int(data["status"]), obj=obj * maxb, obj_scaled=obj,
                                eta=max(data["gap"], data["pinf"], data["dinf"]), hv=data["hv_count"],
                                iters=data["iters"], n=int(mb), m=int(A2.shape[0]),
                                oracle_seconds=time.perf_counter() - t0))


if __name__ == "__main__":
    which = sys.argv[1:] or ["qs60", "bqp60", "bqpsparse", "bqpdual60"]
    for w in which:
        {"qs60": qs60, "bqp60": bqp60, "bqpsparse": bqpsparse, "bqpdual60": bqpdual60}[w]()
