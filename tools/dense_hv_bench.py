"""Hv time of the dense path (BQP moment SDP, ManiSDP_unitdiag closures) vs factor width: FP64 DMMA GEMM efficiency."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manisdp_matlab_b200 import Handle, _lib
from oracle import generators as g
q = int(sys.argv[1]) if len(sys.argv) > 1 else 60
d = np.load(os.path.join(ROOT, "tests", "golden", f"bqp_{q}_1.npz"))
At, b, c, K = g.bqpmom(q, d["Q"], d["e"]); c = c / np.abs(c).max()
b = np.asarray(b.todense()).ravel() if hasattr(b, "todense") else np.asarray(b).ravel()
n = int(K["s"])
with Handle("unitdiag", n, At=At, b=b, c=c) as h:
    h.set_dual(np.zeros(At.shape[1]), 1.0)
    for p in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "8,32,100,200,300,400".split(","))]:
        h.rand_Y(p, 1)
        h.slot_set(_lib.SLOT_U, np.random.default_rng(1).standard_normal((n, p)))
        h.hess_bench(3)
        ms = h.hess_bench(20)
        st = h.stats()
        print(json.dumps(dict(q=q, n=n, p=p, ms=ms, gflops=st.flops_per_hv / ms / 1e6, alg_GBps=st.bytes_per_hv / ms / 1e6)), flush=True)
