"""GPU (needs >= 2 devices, skipped otherwise): the row-sharded handle reproduces the single-GPU solve
(tools/check_sharded.py under torchrun, NCCL).  "er": no locality -> staged peer-copy exchange + column passes;
"torus": locality -> the product gathers the few remote rows directly from the owner's memory."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("graph,port", [("er", "29533"), ("torus", "29534")])
def test_sharded_matches_single_gpu(graph, port):
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tools", "check_sharded.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, CHK_GRAPH=graph))
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    assert json.loads(lines[-1])["sharded_check"] == "ok"


@pytest.mark.parametrize("graph,port", [("er", "29535"), ("torus", "29536")])
def test_column_sharded_matches_single_gpu(graph, port):
    """the column-sharded ("p-sharded") layout under torchrun x2: closures, TR log, point, outer-loop steps"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tools", "check_colsharded.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, CHK_GRAPH=graph))
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    assert json.loads(lines[-1])["colsharded_check"] == "ok"


@pytest.mark.parametrize("p", [6, 18, 40, 64, 100])
def test_column_layout_kernels_on_one_gpu(p):
    """world = 1 column-layout handle: the split closures (raw product + row-sum pass + finishing pass, colshard.cu)
    and the split trust-region loop against the fused single-GPU kernels and the oracle -- the all-reduces are the only
    part of the layout this cannot cover (torchrun test above)."""
    import numpy as np
    from manisdp_matlab_b200 import Handle, problems as P
    from oracle.manisdp_ref import OnlyUnitDiagProblem
    d = np.load(os.path.join(ROOT, "tests", "golden", "G1.npz"))
    C = P.maxcut_C(int(d["n"]), d["ei"].astype(np.int64), d["ej"].astype(np.int64), d["w"].astype(np.float64))
    n = C.shape[0]
    rng = np.random.default_rng(p)
    Y = rng.standard_normal((n, p))
    Y /= np.linalg.norm(Y, axis=1, keepdims=True)
    U = rng.standard_normal((n, p))
    prob = OnlyUnitDiagProblem(C, p, stale_eG=False)
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    with Handle("onlyunitdiag", n, C_csc=C) as h0, Handle("onlyunitdiag", n, C_csc=C, layout="cols") as h:
        for hh in (h0, h):
            hh.set_Y(Y)
        h.col_split()
        assert h.p == p and np.array_equal(h.get_Y(), Y)
        f, f0 = h.cost(), h0.cost()
        assert abs(f - f0) <= 1e-13 * abs(f0) and abs(f - prob.cost(Y)) <= 1e-12 * abs(f)
        (g, gn), (g0, gn0) = h.grad(), h0.grad()
        assert rel(g, g0) < 1e-13 and abs(gn - gn0) <= 1e-13 * gn0 and rel(g, prob.grad(Y)) < 1e-12
        assert rel(h.hess(U), h0.hess(U)) < 1e-13 and rel(h.hess(U), prob.hess(Y, U)) < 1e-12
        i, i0 = h.tr_solve(maxiter=8, maxinner=25, tolgradnorm=1e-9), h0.tr_solve(maxiter=8, maxinner=25, tolgradnorm=1e-9, use_graph=0)
        assert [(r.numinner, r.accepted, r.stop_inner) for r in h.tr_log()] == \
               [(r.numinner, r.accepted, r.stop_inner) for r in h0.tr_log()]
        assert abs(i.cost - i0.cost) <= 1e-10 * abs(i0.cost) and i.hv_count == i0.hv_count
        with pytest.raises(Exception):
            h.kkt(4, 1e-8, 0)  # outer-loop steps need the merged factor
        h.col_merge()
        assert rel(h.get_Y(), h0.get_Y()) < 1e-8
        k, k0 = h.kkt(4, 1e-9, 0), h0.kkt(4, 1e-9, 0)
        assert abs(k.obj - k0.obj) <= 1e-10 * abs(k0.obj) and abs(k.dinf - k0.dinf) <= 1e-3 * k0.dinf + 1e-9
