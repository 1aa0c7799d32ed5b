"""torchrun worker: the row-sharded ONLYUNITDIAG handle (NCCL all-gather of the thin factor + all-reduce of the tCG
scalar packet) must reproduce the single-GPU solve.  Rank 0 also runs the unsharded handle on its own GPU."""
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
    n, deg, p = int(os.environ.get("CHK_N", 20011)), 12, int(os.environ.get("CHK_P", 16))
    if os.environ.get("CHK_GRAPH", "er") == "torus":  # locality: the direct peer-gather product is selected
        n, ei, ej, w = P.synthetic_torus(int(round(n ** 0.5)), seed=3)
    else:
        n, ei, ej, w = P.synthetic_er(n, deg, seed=3)
    C = P.maxcut_C(n, ei, ej, w)
    rng = np.random.default_rng(0)
    Y0 = rng.standard_normal((n, p))
    Y0 /= np.linalg.norm(Y0, axis=1, keepdims=True)
    obj = [_lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    rpr = (n + world - 1) // world
    r0, r1 = min(n, rank * rpr), min(n, (rank + 1) * rpr)
    h = Handle("onlyunitdiag", n, C_csc=C[:, r0:r1], device=local, rank=rank, world=world, row_begin=r0, row_end=r1,
               nccl_id=obj[0])
    h.set_Y(Y0[r0:r1])
    f = h.cost()
    info = h.tr_solve(maxiter=8, maxinner=20, tolgradnorm=1e-9, use_graph=0)
    log = [(r.cost, r.gradnorm, r.numinner, r.accepted, r.stop_inner) for r in h.tr_log()]
    Yloc = h.get_Y()
    k = h.kkt(4, 1e-8, 0)
    r_cut, _ = h.rank_cut(1e-1, apply=False)
    # second phase: a wider factor forces a reallocation of the work arrays (peer mappings are re-published)
    p2 = 3 * p
    Y2 = np.random.default_rng(7).standard_normal((n, p2))
    Y2 /= np.linalg.norm(Y2, axis=1, keepdims=True)
    h.set_Y(Y2[r0:r1])
    info2 = h.tr_solve(maxiter=3, maxinner=10, tolgradnorm=1e-9, use_graph=0)
    h.close()
    parts = [None] * world
    dist.all_gather_object(parts, Yloc)
    ok = True
    if rank == 0:
        Ysh = np.vstack(parts)
        with Handle("onlyunitdiag", n, C_csc=C, device=local) as h1:
            h1.set_Y(Y0)
            f1 = h1.cost()
            info1 = h1.tr_solve(maxiter=8, maxinner=20, tolgradnorm=1e-9, use_graph=0)
            log1 = [(r.cost, r.gradnorm, r.numinner, r.accepted, r.stop_inner) for r in h1.tr_log()]
            Y1 = h1.get_Y()
            k1 = h1.kkt(4, 1e-8, 0)
            r1c, _ = h1.rank_cut(1e-1, apply=False)
            h1.set_Y(Y2)
            info21 = h1.tr_solve(maxiter=3, maxinner=10, tolgradnorm=1e-9, use_graph=0)
        errY = np.linalg.norm(Ysh - Y1) / np.linalg.norm(Y1)
        same_path = [(a[2], a[3], a[4]) for a in log] == [(a[2], a[3], a[4]) for a in log1]
        errc = max(abs(a[0] - b[0]) / abs(b[0]) for a, b in zip(log, log1))
        ok = (abs(f - f1) <= 1e-12 * abs(f1) and same_path and errc < 1e-10 and errY < 1e-7
              and info.hv_count == info1.hv_count and abs(k.dinf - k1.dinf) <= 1e-3 * abs(k1.dinf) + 1e-9
              and abs(k.obj - k1.obj) <= 1e-9 * abs(k1.obj) and r_cut == r1c
              and info2.hv_count == info21.hv_count and abs(info2.cost - info21.cost) <= 1e-10 * abs(info21.cost))
        print(json.dumps({"sharded_check": "ok" if ok else "FAIL", "world": world, "n": n, "p": p, "errY": errY,
                          "err_cost": errc, "same_path": same_path, "hv": [int(info.hv_count), int(info1.hv_count)],
                          "dinf": [k.dinf, k1.dinf], "lam_min": [k.lam_min, k1.lam_min], "rank": [r_cut, r1c]}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
