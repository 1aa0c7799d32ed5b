import time, torch, numpy as np, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n, p = 10**6, 64
h = torch.empty((n, p), dtype=torch.float64).pin_memory()
d = torch.empty((n, p), dtype=torch.float64, device="cuda")
for name, f in [("h2d pinned", lambda: d.copy_(h, non_blocking=True)), ("d2h pinned", lambda: h.copy_(d, non_blocking=True))]:
    f(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): f()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    print(name, "%.1f ms  %.1f GB/s" % (dt * 1e3, n * p * 8 / dt / 1e9))
from manisdp_matlab_b200 import Handle, _lib
import scipy.sparse as sp
with Handle("onlyunitdiag", n, C_csc=sp.identity(n, format="csc")) as hh:
    Y = h.numpy(); Y[:] = 1.0
    hh.set_Y(Y)
    t = time.perf_counter()
    for _ in range(5): hh.set_Y(Y)
    dt = (time.perf_counter() - t) / 5
    print("engine set_Y", "%.1f ms  %.1f GB/s" % (dt * 1e3, n * p * 8 / dt / 1e9))
    out = torch.empty((n, p), dtype=torch.float64).pin_memory().numpy()
    t = time.perf_counter()
    for _ in range(5): hh.lib.manisdp_get_Y(hh._h, _lib._pf(out), 0)
    dt = (time.perf_counter() - t) / 5
    print("engine get_Y", "%.1f ms  %.1f GB/s" % (dt * 1e3, n * p * 8 / dt / 1e9))
