"""Hessian-product cost of a dual (ManiDSDP) handle on the SOS form of BQP q at width p:
    python tools/dual_hv_bench.py [q] [p] [reps]        (measurement script)"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instances import generators as G  # noqa: E402
from manisdp_matlab_b200 import Handle  # noqa: E402

q, p, reps = (int(v) for v in (sys.argv[1:4] + ["60", "64", "20"][len(sys.argv) - 1:]))
d = np.load(os.path.join(ROOT, "tests", "golden", f"bqp_{q}_1.npz"))
A, b, dAAt, mb = G.bqpsos(d["Q"], d["e"], q)
rng = np.random.default_rng(0)
with Handle("dual_unitdiag", mb, At=A.T.tocsc(), b=b / np.abs(b).max(), c=np.zeros(mb * mb), dAAt=dAAt) as h:
    h.set_sigma(0.01)
    h.rand_Y(p, 1)
    h.cost()
    h.slot_set(7, h.project(rng.standard_normal((mb, p))))
    h.hess_bench(3)
    ms = h.hess_bench(reps)
    st = h.stats()
    print(json.dumps(dict(q=q, n=mb, m=A.shape[0], p=p, ms_per_hv=ms, flops_per_hv=st.flops_per_hv,
                          tflops=st.flops_per_hv / (ms * 1e-3) / 1e12, modes=[st.s_mode, st.a_mode])))
