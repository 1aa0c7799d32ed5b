"""torchrun worker: the COLUMN-sharded ONLYUNITDIAG handle (every rank: all rows, ceil(p/G) columns of the factor; NCCL
all-reduces of per-row scalars, no exchange of the factor -- csrc/colshard.cu) must reproduce the single-GPU solve:
closures, trust-region log, final point, and the outer-loop steps after the merge.  Rank 0 runs the unsharded handle."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from manisdp_matlab_b200 import Handle, _lib, problems as P

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, deg, p = int(os.environ.get("CHK_N", 20011)), 12, int(os.environ.get("CHK_P", 18))
    UG = int(os.environ.get("CHK_USE_GRAPH", 0))  # 1 with MANISDP_COL_GRAPH=1: the split loop as a CUDA graph
    if os.environ.get("CHK_GRAPH", "er") == "torus":
        n, ei, ej, w = P.synthetic_torus(int(round(n ** 0.5)), seed=3)
    else:
        n, ei, ej, w = P.synthetic_er(n, deg, seed=3)
    C = P.maxcut_C(n, ei, ej, w)
    rng = np.random.default_rng(0)
    Y0 = rng.standard_normal((n, p))
    Y0 /= np.linalg.norm(Y0, axis=1, keepdims=True)
    U = rng.standard_normal((n, p))
    obj = [_lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    pl = (p + world - 1) // world
    pfull = pl * world
    Y0p = np.hstack([Y0, np.zeros((n, pfull - p))])  # the split pads the factor to a multiple of the world size
    Up = np.hstack([U, np.zeros((n, pfull - p))])
    h = Handle("onlyunitdiag", n, C_csc=C, device=local, rank=rank, world=world, nccl_id=obj[0], layout="cols")
    h.set_Y(Y0)
    h.col_split()
    assert h.p == pl
    sl = slice(rank * pl, (rank + 1) * pl)
    assert np.array_equal(h.get_Y(), Y0p[:, sl])
    f = h.cost()
    g, gn = h.grad()
    hv = h.hess(Up[:, sl])
    info = h.tr_solve(maxiter=8, maxinner=20, tolgradnorm=1e-9, use_graph=UG)
    log = [(r.cost, r.gradnorm, r.numinner, r.accepted, r.stop_inner) for r in h.tr_log()]
    h.col_merge()
    assert h.p == pfull
    Ym = h.get_Y()
    k = h.kkt(4, 1e-8, 0)
    r_cut, _ = h.rank_cut(1e-1, apply=False)
    h.escape(max(1, min(int(k.nneg), 4)), 0.5, 0)
    h.col_split()  # second phase: a wider factor (reallocation) through the same handle
    info2 = h.tr_solve(maxiter=3, maxinner=10, tolgradnorm=1e-9, use_graph=UG)
    h.col_merge()
    Ym2 = h.get_Y()
    h.close()
    parts = [None] * world
    dist.all_gather_object(parts, (g, hv, Ym, Ym2, f, gn, int(info.hv_count), float(k.dinf), float(info2.cost)))
    ok = True
    if rank == 0:
        gsh = np.hstack([q[0] for q in parts])
        hsh = np.hstack([q[1] for q in parts])
        same_everywhere = all(np.array_equal(q[2], parts[0][2]) and np.array_equal(q[3], parts[0][3]) and q[4:] == parts[0][4:]
                              for q in parts)
        with Handle("onlyunitdiag", n, C_csc=C, device=local) as h1:
            h1.set_Y(Y0p)
            f1 = h1.cost()
            g1, gn1 = h1.grad()
            hv1 = h1.hess(Up)
            info1 = h1.tr_solve(maxiter=8, maxinner=20, tolgradnorm=1e-9, use_graph=UG)
            log1 = [(r.cost, r.gradnorm, r.numinner, r.accepted, r.stop_inner) for r in h1.tr_log()]
            Y1 = h1.get_Y()
            k1 = h1.kkt(4, 1e-8, 0)
            r1c, _ = h1.rank_cut(1e-1, apply=False)
            h1.escape(max(1, min(int(k1.nneg), 4)), 0.5, 0)
            p_esc = h1.p
            info21 = h1.tr_solve(maxiter=3, maxinner=10, tolgradnorm=1e-9, use_graph=UG)
        rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
        same_path = [(a[2], a[3], a[4]) for a in log] == [(a[2], a[3], a[4]) for a in log1]
        errc = max(abs(a[0] - b[0]) / abs(b[0]) for a, b in zip(log, log1))
        checks = dict(cost=abs(f - f1) <= 1e-12 * abs(f1), grad=rel(gsh, g1) < 1e-12, gradnorm=abs(gn - gn1) <= 1e-12 * gn1,
                      hess=rel(hsh, hv1) < 1e-12, same_path=same_path, log_cost=errc < 1e-10, Y=rel(Ym, Y1) < 1e-7,
                      hv=info.hv_count == info1.hv_count, dinf=abs(k.dinf - k1.dinf) <= 1e-3 * abs(k1.dinf) + 1e-9,
                      obj=abs(k.obj - k1.obj) <= 1e-9 * abs(k1.obj), rank=r_cut == r1c, ranks_agree=same_everywhere,
                      phase2=abs(info2.cost - info21.cost) <= 1e-6 * abs(info21.cost))
        ok = all(checks.values())
        print(json.dumps({"colsharded_check": "ok" if ok else "FAIL", "world": world, "n": n, "p": p, "pl": pl,
                          "checks": {k_: bool(v) for k_, v in checks.items()}, "err_cost": errc,
                          "hv": [int(info.hv_count), int(info1.hv_count)], "dinf": [k.dinf, k1.dinf],
                          "phase2_cost": [info2.cost, info21.cost], "p_after_escape": [int(Ym2.shape[1]), int(p_esc)]}),
              flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
