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


def _g1():
    import numpy as np
    from manisdp_matlab_b200 import problems as P
    d = np.load(os.path.join(ROOT, "tests", "golden", "G1.npz"))
    return P.maxcut_C(int(d["n"]), d["ei"].astype(np.int64), d["ej"].astype(np.int64), d["w"].astype(np.float64))


@pytest.mark.parametrize("ndev", [1, 2])
def test_single_process_group_matches_single_handle(ndev):
    """manisdp_group_* (csrc/group.cu): ONE caller thread drives `ndev` GPUs through worker threads inside the library
    (the mode the MATLAB gateway uses).  Closures, a trust-region solve and the outer-loop steps against a plain handle;
    ndev = 1 runs everywhere, ndev = 2 needs two devices."""
    import numpy as np
    import torch
    from manisdp_matlab_b200 import GroupHandle, Handle
    if torch.cuda.device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    C = _g1()
    n, p = C.shape[0], 16
    rng = np.random.default_rng(4)
    Y = rng.standard_normal((n, p))
    Y /= np.linalg.norm(Y, axis=1, keepdims=True)
    with Handle("onlyunitdiag", n, C_csc=C) as h0, GroupHandle(n, C, list(range(ndev))) as g:
        h0.set_Y(Y)
        g.set_Y(Y)
        assert abs(g.cost() - h0.cost()) <= 1e-13 * abs(h0.cost())
        i0 = h0.tr_solve(maxiter=8, maxinner=25, tolgradnorm=1e-9, use_graph=0)
        ig = g.tr_solve(maxiter=8, maxinner=25, tolgradnorm=1e-9, use_graph=0)
        assert ig.hv_count == i0.hv_count and ig.iters == i0.iters
        assert abs(ig.cost - i0.cost) <= 1e-10 * abs(i0.cost)
        Yg, Y0 = g.get_Y(), h0.get_Y()
        assert Yg.shape == Y0.shape and np.linalg.norm(Yg - Y0) <= 1e-8 * np.linalg.norm(Y0)
        k0, kg = h0.kkt(8, 1e-9, 0), g.kkt(8, 1e-9, 0)
        assert abs(kg.obj - k0.obj) <= 1e-10 * abs(k0.obj) and abs(kg.dinf - k0.dinf) <= 1e-3 * k0.dinf + 1e-9
        assert g.rank_cut(1e-1, apply=False)[0] == h0.rank_cut(1e-1, apply=False)[0]
        g.escape(max(1, min(int(kg.nneg), 8)), 0.5, 0)
        assert g.p == p + max(1, min(int(kg.nneg), 8))


@pytest.mark.parametrize("ndev", [1, 2])
def test_driver_and_mex_gateway_on_a_device_group(ndev):
    """the drop-in ManiSDP_onlyunitdiag with options.devices, and the same solve through mexFunction with a device list
    (matlab/manisdp_mex.cpp 'create' with a 5th argument): G1 to dinf <= 1e-8 at the known optimum"""
    import numpy as np
    import torch
    from manisdp_matlab_b200 import ManiSDP_onlyunitdiag
    if torch.cuda.device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    C = _g1()
    X, obj, data = ManiSDP_onlyunitdiag(C, dict(p0=40, verbose=False, devices=list(range(ndev))))
    assert data["dinf"] < 1e-8 and data["n_devices"] == ndev
    assert abs(obj - (-12083.19765455)) <= 1e-6 * 12083.2
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_mex_gateway import Mex
    mex = Mex()
    n = C.shape[0]
    (hd,) = mex.call(1, "create", 0.0, float(n), C, np.arange(ndev, dtype=np.float64).reshape(1, -1))
    mex.call(0, "rand_Y", None, 40.0, 0.0, handle=hd)
    (s,) = mex.call(1, "tr_solve", None, dict(maxiter=40, maxinner=100, tolgradnorm=1e-8, use_graph=0), handle=hd)
    (kk,) = mex.call(1, "kkt", None, 8.0, -1e-8, 0.0, handle=hd)
    assert mex.field(kk, "dinf") < 1e-8 and abs(mex.field(kk, "obj") - (-12083.19765455)) <= 1e-6 * 12083.2
    (st,) = mex.call(1, "stats", None, handle=hd)
    assert mex.field(st, "n_devices") == ndev and mex.field(st, "p") == 40
    (Yh,) = mex.call(1, "get_Y", None, 0.0, handle=hd)
    assert mex.mat(Yh).shape == (40, n)
    with pytest.raises(RuntimeError, match="not available on a multi-GPU group"):
        mex.call(0, "hess", None, np.zeros((40, n)), handle=hd)
    mex.call(0, "destroy", None, handle=hd)
    mex.lib.stub_free(hd)
