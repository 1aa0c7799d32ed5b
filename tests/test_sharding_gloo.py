"""CPU, world_size 2 over gloo: the host-side logic of the row-sharded path -- the partition, the shard of C, and the
exchange protocol (all-gather of the thin factor before the product, all-reduce of the scalar packet after it) --
reproduces the unsharded closures.  The arithmetic here is NumPy standing in for the kernels; the CUDA path itself is
covered by tests/test_gpu_sharded.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, p, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from manisdp_matlab_b200 import problems as P, sharding as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nn, ei, ej, w = P.synthetic_er(n, 6, seed=1)
    C = P.maxcut_C(nn, ei, ej, w)
    rng = np.random.default_rng(0)
    Y = rng.standard_normal((n, p))
    Y /= np.linalg.norm(Y, axis=1, keepdims=True)
    U = rng.standard_normal((n, p))
    Cs, r0, r1 = S.shard_C(C, world, rank)
    rpr = S.rows_per_rank(n, world)
    # exchange step: equal-count all-gather with zero padding (dist.h layout)
    def allgather_rows(Aloc):
        pad = np.zeros((rpr, p))
        pad[: r1 - r0] = Aloc
        outs = [torch.zeros(rpr, p, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(outs, torch.from_numpy(pad))
        return torch.cat(outs).numpy()[:n] if False else torch.cat(outs).numpy()
    Yfull = allgather_rows(Y[r0:r1])
    assert np.array_equal(Yfull[:n], Y)
    Cpad = Cs.T.tocsr()  # owned rows x n
    # the gathered array has world*rpr rows (padding rows are zero and never referenced by a column index < n)
    YC = Cpad @ Yfull[:n]
    eG = np.sum(YC * Y[r0:r1], axis=1, keepdims=True)
    pkt = torch.tensor([eG.sum(), np.sum((YC - Y[r0:r1] * eG) ** 2)], dtype=torch.float64)
    dist.all_reduce(pkt)
    Ufull = allgather_rows(U[r0:r1])[:n]
    eH = Cpad @ Ufull
    H = eH - Y[r0:r1] * np.sum(Y[r0:r1] * eH, axis=1, keepdims=True) - U[r0:r1] * eG
    dHd = torch.tensor([np.sum(U[r0:r1] * H)], dtype=torch.float64)
    dist.all_reduce(dHd)
    # unsharded reference
    YCf = C.T @ Y
    eGf = np.sum(YCf * Y, axis=1, keepdims=True)
    eHf = C.T @ U
    Hf = eHf - Y * np.sum(Y * eHf, axis=1, keepdims=True) - U * eGf
    ok = (abs(pkt[0].item() - eGf.sum()) <= 1e-12 * abs(eGf.sum())
          and abs(pkt[1].item() - np.sum((YCf - Y * eGf) ** 2)) <= 1e-12 * np.sum((YCf - Y * eGf) ** 2)
          and np.allclose(H, Hf[r0:r1], rtol=0, atol=1e-13)
          and abs(dHd.item() - np.sum(U * Hf)) <= 1e-12 * abs(np.sum(U * Hf)))
    q.put((rank, bool(ok), r0, r1))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [101, 128])
def test_row_sharded_protocol_world2(n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + n % 50
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 5, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
    assert all(r[1] for r in res), res
    # the partition tiles [0, n) without overlap
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == n


def test_partition_properties():
    from manisdp_matlab_b200 import sharding as S
    for n in [1, 7, 8, 1000, 10**6 + 3]:
        for world in [1, 2, 3, 4, 8]:
            cover = 0
            prev = 0
            for r in range(world):
                a, b = S.row_range(n, world, r)
                assert a == min(n, prev) and a <= b <= n
                assert b - a <= S.rows_per_rank(n, world)
                cover += b - a
                prev = b
            assert cover == n
