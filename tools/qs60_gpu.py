"""Config 3 (quartic on the sphere, q = 60, N(0,1) coefficients of default_rng(seed)) through the drop-in ManiSDP.
    python tools/qs60_gpu.py <coefficient seed> ['{"delta": 6, ...}'] [--verbose]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instances import generators as g
from manisdp_matlab_b200 import ManiSDP

seed = int(sys.argv[1])
extra = json.loads(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].startswith("{") else {}
coe = np.random.default_rng(seed).standard_normal(635376)
t0 = time.perf_counter()
At, b, c, K = g.qsmom(60, coe)
tg = time.perf_counter() - t0
o = dict(tol=1e-8, theta=1e-2, tau1=0.02, verbose="--verbose" in sys.argv)
o.update(extra)
t0 = time.perf_counter()
X, obj, data = ManiSDP(At, np.asarray(b.todense()).ravel() if hasattr(b, "todense") else b, c, K, o)
print(json.dumps(dict(config="qs60", coe_seed=seed, options=extra, obj=obj, eta=max(data["gap"], data["pinf"], data["dinf"]),
                      iters=data["iters"], hv=data["hv_count"], status=data["status"], seconds=time.perf_counter() - t0,
                      tr_seconds=data["tr_seconds"], gen_seconds=tg, phase_seconds=data.get("phase_seconds"), eig_iters=data.get("eig_iters_total"), p_max=max(data["fac_size"]))), flush=True)
