"""Where does the wall time of a drop-in solve go?  Wraps the Handle methods with timers."""
import json, os, sys, time
from collections import defaultdict
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import manisdp_matlab_b200 as M
from manisdp_matlab_b200 import _lib
from oracle import generators as g

T = defaultdict(float); N = defaultdict(int); EXTRA = defaultdict(list)
def wrap(name):
    f = getattr(_lib.Handle, name)
    def w(self, *a, **k):
        t = time.perf_counter(); r = f(self, *a, **k); T[name] += time.perf_counter() - t; N[name] += 1
        if name == "kkt": EXTRA["eig_iters"].append(r.eig_iters); EXTRA["eig_resid"].append(r.eig_resid)
        return r
    setattr(_lib.Handle, name, w)
for nm in ["__init__", "tr_solve", "kkt", "rank_cut", "escape", "line_search", "get_Y", "set_dual", "rand_Y"]:
    wrap(nm)

name = sys.argv[1] if len(sys.argv) > 1 else "bqp60"
q = int(name[3:])
d = np.load(os.path.join(ROOT, "tests", "golden", f"bqp_{q}_1.npz"))
At, b, c, K = g.bqpmom(q, d["Q"], d["e"]); c = c / np.abs(c).max()
b = np.asarray(b.todense()).ravel() if hasattr(b, "todense") else np.asarray(b).ravel()
t = time.perf_counter()
X, obj, data = M.ManiSDP_unitdiag(At, b, c, K, dict(verbose=False, eig_tol=float(sys.argv[2]) if len(sys.argv) > 2 else 0.0))
tot = time.perf_counter() - t
print(json.dumps(dict(config=name, total=tot, obj=obj, iters=data["iters"], hv=int(data["hv_count"]), tr_dev=data["tr_seconds"],
      times={k: round(v, 3) for k, v in T.items()}, calls=dict(N), eig_iters=EXTRA["eig_iters"], p=data["fac_size"],
      eig_resid_max=max(EXTRA["eig_resid"]))))
