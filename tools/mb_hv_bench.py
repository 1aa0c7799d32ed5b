"""Hessian-product cost of a multi-block handle on the sparse-BQP example (example/example_bqp_sparse.m, t cliques of q
variables) at a given common width p:  python tools/mb_hv_bench.py [t] [q] [p] [reps]   (measurement script)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instances import generators as G  # noqa: E402
from manisdp_matlab_b200 import Handle  # noqa: E402

t, q, p, reps = (int(v) for v in (sys.argv[1:5] + ["20", "20", "100", "20"][len(sys.argv) - 1:]))
At, b, c, K, n, I, coe = G.bqp_sparse_instance(t, q, 1)
ns = K["s"]
rng = np.random.default_rng(0)
with Handle("multiblock", sum(ns), At=At, b=b, c=c, block_sizes=ns, nob=K["nob"]) as h:
    h.set_dual(0.01 * rng.standard_normal(At.shape[1]), 0.5)
    h.mb_rand_Y([min(p, v) for v in ns], 1)
    h.cost()
    h.slot_set(7, h.project(rng.standard_normal((h.n, h.p))))
    h.hess_bench(3)
    ms = h.hess_bench(reps)
    t0 = time.perf_counter()
    info = h.tr_solve(maxiter=2, maxinner=20, tolgradnorm=1e-12)
    dt = time.perf_counter() - t0
    st = h.stats()
    print(json.dumps(dict(t=t, q=q, p=p, N=int(sum(ns)), m=At.shape[1], nnzA=int(At.nnz), ms_per_hv=ms,
                          tr_hv=int(info.hv_count), tr_ms_per_hv=1e3 * info.seconds / max(1, info.hv_count), tr_wall=dt,
                          bytes_per_hv=st.bytes_per_hv)))
