"""Time-to-KKT of BASELINE.json configs 1-4 through the drop-in drivers (measurement script, not product code: the
SeDuMi inputs are built with the oracle's restatement of the reference's generators bqpmom / qsmom / generate_hamming).

    python tools/run_configs.py [g1 g11 g32 bqp20 bqp60 qs20 qs40s theta98 theta102 er:100000 torus:316] [--cpu] [--verbose]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

import manisdp_matlab_b200 as M  # noqa: E402
from manisdp_matlab_b200 import problems as P  # noqa: E402
from oracle import generators as g  # noqa: E402
from oracle import manisdp_ref as ref  # noqa: E402


def dense_b(b):
    return np.asarray(b.todense()).ravel() if hasattr(b, "todense") else np.asarray(b).ravel()


def run(name, cpu):
    t0 = time.perf_counter()
    if name in ("g1", "g11", "g32"):
        d = np.load(os.path.join(GOLDEN, f"{name.upper()}.npz"))
        C = P.maxcut_C(int(d["n"]), d["ei"].astype(np.int64), d["ej"].astype(np.int64), d["w"].astype(np.float64))
        opts = dict(p0=40)
        call = lambda mod, o: mod.ManiSDP_onlyunitdiag(C, o)
        n, m = C.shape[0], C.shape[0]
    elif name.startswith("er:") or name.startswith("torus:"):  # config 5 family: er:<n> (mean degree 48), torus:<side>
        arg = int(float(name.split(":")[1]))
        if name.startswith("er:"):
            nn, ei, ej, w = P.synthetic_er(arg, 48, seed=0)
        else:
            nn, ei, ej, w = P.synthetic_torus(arg, seed=0)
        C = P.maxcut_C(nn, ei, ej, w)
        opts = dict(p0=64, delta=8)
        call = lambda mod, o: mod.ManiSDP_onlyunitdiag(C, o)
        n, m = nn, nn
    elif name.startswith("bqp") and name[3:].isdigit():
        q = int(name[3:])
        d = np.load(os.path.join(GOLDEN, f"bqp_{q}_1.npz"))
        At, b, c, K = g.bqpmom(q, d["Q"], d["e"])
        c = c / np.abs(c).max()
        b = dense_b(b)
        opts = dict(tol=1e-8)
        call = lambda mod, o: mod.ManiSDP_unitdiag(At, b, c, K, o)
        n, m = int(K["s"]), At.shape[1]
    elif name.startswith("qs"):
        q = int(name[2:4])
        if name.endswith("s"):  # synthetic coefficients (qs_c_60 is missing from the reference tree)
            from math import comb
            coe = np.random.default_rng(q).standard_normal(comb(q + 4, 4) if False else g.qs_num_coe(q))
        else:
            coe = np.load(os.path.join(GOLDEN, f"qs_c_{q}_1.npz"))["coe"]
        At, b, c, K = g.qsmom(q, coe)
        b = dense_b(b)
        opts = dict(tol=1e-8, theta=1e-2, tau1=0.02)
        call = lambda mod, o: mod.ManiSDP(At, b, c, K, o)
        n, m = int(K["s"]), At.shape[1]
    elif name.startswith("theta"):
        k, dd = {"theta756": (7, [5, 6]), "theta98": (9, 8), "theta102": (10, 2), "theta112": (11, 2)}[name]
        At, b, c, K = g.generate_hamming(k, dd)
        b = dense_b(b)
        opts = dict(tol=1e-8, sigma0=1e5, sigma_max=1e8, line_search=1, TR_maxiter=10, TR_maxinner=100)
        call = lambda mod, o: mod.ManiSDP_unittrace(At, b, c, K, o)
        n, m = int(K["s"]), At.shape[1]
    elif name.startswith("bqpsparse:"):  # multi-block: example/example_bqp_sparse.m (t cliques of q variables)
        from instances import generators as gi
        t_, q_ = (int(v) for v in name.split(":")[1].split("x"))
        At, b, c, K, _, _, _ = gi.bqp_sparse_instance(t_, q_, 1)
        opts = dict(tol=1e-8, line_search=1, tau1=1)
        call = lambda mod, o: mod.ManiSDP_multiblock(At, b, c, K, o)
        n, m = int(sum(K["s"])), At.shape[1]
    elif name.startswith("bqpdual"):  # dual approach: example/dual/example_bqp_dual.m on data/bqp_{Q,e}_<q>_1.txt
        import scipy.sparse as sp_
        from instances import generators as gi
        q = int(name[len("bqpdual"):])
        fn = os.path.join(GOLDEN, f"bqp_{q}_1.npz")
        if os.path.exists(fn):
            d = np.load(fn)
        else:  # beyond the reference's data files (d <= 60): example_bqp_dual.m:2-5 with NumPy's generator, seed = q
            rq = np.random.default_rng(q)
            Qs = rq.standard_normal((q, q))
            d = {"Q": (Qs + Qs.T) / 2, "e": rq.standard_normal(q)}
        A, bb, dAAt, mb = gi.bqpsos(d["Q"], d["e"], q)
        v = np.zeros((A.shape[0], 1))
        v[0] = 1.0
        A2 = sp_.hstack([sp_.csr_matrix(v), A]).tocsr()
        c = np.concatenate([[1.0], np.zeros(mb * mb)])
        maxb = float(np.abs(bb).max())
        b, K = bb / maxb, {"f": 1, "s": mb}
        opts = dict(dAAt=dAAt, tol=1e-8, line_search=1)

        def call(mod, o, _A=A2, _b=b, _c=c, _K=K, _s=maxb):
            X_, obj_, data_ = mod.ManiDSDP_unitdiag(_A, _b, _c, _K, o)
            return X_, obj_ * _s, data_
        n, m = mb, A2.shape[0]
    elif name == "demo1":  # multi-block: data/test.m on data/SDP_demo_1.mat (89 blocks, K.nob = 0)
        import scipy.sparse as sp_
        d = np.load(os.path.join(GOLDEN, "sdp_demo_1.npz"))
        At = sp_.csc_matrix((d["At_data"], d["At_indices"], d["At_indptr"]), shape=tuple(d["At_shape"]))
        c = np.zeros(At.shape[0])
        c[d["c_idx"]] = d["c_val"]
        b, K = d["b"], {"s": [int(v) for v in d["ns"]], "nob": 0}
        opts = dict(tol=1e-4, gama=2, alpha=0.1, sigma0=1e-2, TR_maxinner=50, TR_maxiter=50, theta=1e-3, delta=4,
                    line_search=0, AL_maxiter=300)
        call = lambda mod, o: mod.ManiSDP_multiblock(At, b, c, K, o)
        n, m = int(sum(K["s"])), At.shape[1]
    else:
        raise SystemExit(f"unknown config {name}")
    t_gen = time.perf_counter() - t0
    o = dict(opts, verbose="--verbose" in sys.argv)
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:  # torchrun: config-5 family on a column-sharded handle, one rank per GPU
        import torch
        import torch.distributed as dist
        rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        obj = [M._lib.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        o.update(world=world, rank=rank, device=local, nccl_id=obj[0], verbose=o["verbose"] and rank == 0)
    for a in sys.argv[1:]:
        if a.startswith("--opts="):  # option overrides as JSON, e.g. --opts='{"p0": 256, "delta": 24}'
            o.update(json.loads(a[len("--opts="):]))
    t0 = time.perf_counter()
    X, obj, data = call(M, o)
    dt = time.perf_counter() - t0
    eta = max(data.get("gap", 0.0), data.get("pinf", 0.0), data["dinf"])
    rec = dict(config=name, n=n, m=m, seconds=dt, obj=obj, eta=eta, iters=data["iters"], hv=int(data["hv_count"]),
               tr_seconds=data["tr_seconds"], hv_per_s=data["hv_count"] / max(data["tr_seconds"], 1e-9),
               status=data["status"], launches=int(data.get("launches", 0)), gen_seconds=t_gen,
               modes=[data.get("s_mode"), data.get("a_mode")], p_max=int(np.max([np.max(v) for v in data["fac_size"]])),
               kkt_seconds=data.get("kkt_seconds"), eig_iters=data.get("eig_iters_total"),
               setup_seconds=data.get("setup_seconds"), fac_size=data["fac_size"] if np.ndim(data["fac_size"]) == 1 else [max(v) for v in data["fac_size"]],
               phase_seconds=data.get("phase_seconds"),
               options={k: v for k, v in o.items() if k not in ("verbose", "nccl_id", "dAAt")}, n_gpus=world)
    if world > 1 and int(os.environ["RANK"]) != 0:
        return
    if cpu:
        t0 = time.perf_counter()
        Xc, objc, dc = call(ref, dict(opts, seed=0))
        rec.update(cpu_seconds=time.perf_counter() - t0, cpu_obj=objc, cpu_hv=int(dc["hv_count"]),
                   cpu_iters=dc["iters"], cpu_eta=max(dc.get("gap", 0.0), dc.get("pinf", 0.0), dc["dinf"]),
                   cpu_cores=os.cpu_count())
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["g1", "g11", "bqp20", "theta756"]
    for nm in names:
        try:
            run(nm, "--cpu" in sys.argv)
        except Exception as e:  # keep going: one failing config must not hide the others
            print(json.dumps(dict(config=nm, error=repr(e))), flush=True)
