"""Hv-kernel time vs factor width on the n=1e6 synthetic MaxCut graphs (gather-locality experiment)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manisdp_matlab_b200 import Handle, _lib, problems as P

def run(name, C, n, ps, reps):
    out = []
    with Handle("onlyunitdiag", n, C_csc=C) as h:
        for p in ps:
            h.rand_Y(p, 1)
            U = np.random.default_rng(1).standard_normal((n, p))
            h.slot_set(_lib.SLOT_U, U)
            h.hess_bench(40)  # long enough for the SM clock to ramp after the host-side set-up above
            ms = h.hess_bench(reps)
            st = h.stats()
            out.append(dict(graph=name, p=p, narrow=os.environ.get("MANISDP_SPMM_NARROW", "1"), ms=ms, alg_GBps=st.bytes_per_hv / ms / 1e6,
                            gather_GBps=(st.nnzC * p * 8) / ms / 1e6))
            print(json.dumps(out[-1]), flush=True)
    return out

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "er"
    ps = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 8, 16, 32, 64, 128]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 10**6
    if which == "er":
        n, ei, ej, w = P.synthetic_er(n, 48, 0)
    else:
        n, ei, ej, w = P.synthetic_torus(int(round(n ** 0.5)), 0)
    C = P.maxcut_C(n, ei, ej, w)
    if len(sys.argv) > 4 and sys.argv[4] == "narrow_ab":  # generic kernel vs k_spmm_narrow on the same graph
        os.environ["MANISDP_SPMM_NARROW"] = "0"
        run(which, C, n, ps, 20)
        os.environ["MANISDP_SPMM_NARROW"] = "1"
    run(which, C, n, ps, 20)
