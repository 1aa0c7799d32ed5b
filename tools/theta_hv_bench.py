"""Hv time of the ManiSDP_unittrace closures on theta of Hamming(k,d) (sparse A path: K2 SDDMM gather + K3 scatter-SpMM
around one dense n x n x p product with the dual slack).    python tools/theta_hv_bench.py [k] [d] [p,p,...] [reps]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manisdp_matlab_b200 import Handle, _lib
from instances import generators as g
k = int(sys.argv[1]) if len(sys.argv) > 1 else 11
d = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ps = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "8,20,32").split(",")]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 50
At, b, c, K = g.generate_hamming(k, d)
b = np.asarray(b.todense()).ravel() if hasattr(b, "todense") else np.asarray(b).ravel()
n = int(K["s"])
with Handle("unittrace", n, At=At, b=b, c=c) as h:
    h.set_dual(np.zeros(At.shape[1]), 1e5)
    for p in ps:
        h.rand_Y(p, 1)
        h.slot_set(_lib.SLOT_U, np.random.default_rng(1).standard_normal((n, p)))
        h.hess_bench(5)
        ms = h.hess_bench(reps)
        st = h.stats()
        print(json.dumps(dict(k=k, d=d, n=n, m=int(At.shape[1]), p=p, ms=ms, alg_GBps=st.bytes_per_hv / ms / 1e6,
                              modes=[int(st.s_mode), int(st.a_mode)])), flush=True)
