"""Generates the committed fixtures under tests/golden/ (run in the BUILD container, where /root/reference exists):

  * instance data copied out of the reference's data/ directory as compact .npz files (G-set edge lists, BQP
    coefficient files, quartic-sphere coefficient files) -- the GPU box has no /root/reference;
  * oracle outputs on those instances: optima / KKT residues of complete solves and per-iteration trust-region logs
    from fixed starting points (for iterate-level parity of the device RTR loop).

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/data"

from oracle import generators as g  # noqa: E402
from oracle import manisdp_ref as ref  # noqa: E402
from oracle.manopt_rtr import trustregions  # noqa: E402


def save_instances():
    for name in ["G1", "G11", "G32"]:
        n, ei, ej, w = g.read_gset(f"{REF}/Gset/{name}.txt")
        np.savez_compressed(f"{HERE}/{name}.npz", n=n, ei=ei.astype(np.int32), ej=ej.astype(np.int32),
                            w=w.astype(np.int8))
    for q in [10, 20, 60]:
        Q = np.loadtxt(f"{REF}/bqp_Q_{q}_1.txt", delimiter=",")
        e = np.loadtxt(f"{REF}/bqp_e_{q}_1.txt")
        np.savez_compressed(f"{HERE}/bqp_{q}_1.npz", Q=Q, e=e)
    for q in [10, 20]:
        coe = np.loadtxt(f"{REF}/qs_c_{q}_1.txt")
        np.savez_compressed(f"{HERE}/qs_c_{q}_1.npz", coe=coe)


def tr_log(res):
    return [dict(iter=i.iter, cost=i.cost, gradnorm=i.gradnorm, Delta=i.Delta, rho=(i.rho if np.isfinite(i.rho) else None),
                 accepted=bool(i.accepted), numinner=i.numinner, stop_inner=i.stop_inner) for i in res.info]


def main():
    save_instances()
    out = {}
    # ---- MaxCut G1: iterate-level log from a fixed start + full solve
    d = np.load(f"{HERE}/G1.npz")
    C = g.maxcut_C(int(d["n"]), d["ei"].astype(np.int64), d["ej"].astype(np.int64), d["w"].astype(np.float64))
    rng = np.random.default_rng(123)
    Y0 = rng.standard_normal((C.shape[0], 12))
    Y0 /= np.linalg.norm(Y0, axis=1, keepdims=True)
    prob = ref.OnlyUnitDiagProblem(C, 12, stale_eG=False)
    res = trustregions(prob, Y0.copy(), maxiter=12, maxinner=30, tolgradnorm=1e-8)
    out["G1_tr_log_seed123_p12"] = tr_log(res)
    X, obj, data = ref.ManiSDP_onlyunitdiag(C, dict(p0=40, seed=0))
    out["G1_opt"] = dict(obj=obj, dinf=data["dinf"], hv=data["hv_count"], iters=data["iters"])
    for name in ["G11"]:
        d = np.load(f"{HERE}/{name}.npz")
        Cn = g.maxcut_C(int(d["n"]), d["ei"].astype(np.int64), d["ej"].astype(np.int64), d["w"].astype(np.float64))
        X, obj, data = ref.ManiSDP_onlyunitdiag(Cn, dict(p0=40, seed=0))
        out[f"{name}_opt"] = dict(obj=obj, dinf=data["dinf"], hv=data["hv_count"], iters=data["iters"])
    # ---- BQP 10 / 20 through ManiSDP_unitdiag
    for q in [10, 20]:
        d = np.load(f"{HERE}/bqp_{q}_1.npz")
        At, b, c, K = g.bqpmom(q, d["Q"], d["e"])
        mc = float(np.abs(c).max())
        X, obj, data = ref.ManiSDP_unitdiag(At, b, c / mc, K, dict(seed=0))
        out[f"bqp_{q}_1_opt"] = dict(obj_scaled=obj, obj=obj * mc, eta=max(data["gap"], data["pinf"], data["dinf"]),
                                     hv=data["hv_count"], iters=data["iters"], n=int(K["s"]), m=int(At.shape[1]))
        if q == 10:
            out["bqp_10_1_bruteforce"] = g.bqp_bruteforce(d["Q"], d["e"])
    # ---- quartic sphere 10 through ManiSDP (general)
    d = np.load(f"{HERE}/qs_c_10_1.npz")
    At, b, c, K = g.qsmom(10, d["coe"])
    X, obj, data = ref.ManiSDP(At, b, c, K, dict(seed=0, tol=1e-8, theta=1e-2, tau1=0.02))
    out["qs_c_10_1_opt"] = dict(obj=obj, eta=max(data["gap"], data["pinf"], data["dinf"]), hv=data["hv_count"],
                                iters=data["iters"], n=int(K["s"]), m=int(At.shape[1]))
    # ---- theta of Hamming(7,[5,6]) through ManiSDP_unittrace (example options)
    At, b, c, K = g.generate_hamming(7, [5, 6])
    X, obj, data = ref.ManiSDP_unittrace(At, b, c, K, dict(seed=0, tol=1e-6, sigma0=1e5, sigma_max=1e8, line_search=1))
    out["hamming_7_5_6_opt"] = dict(obj=obj, eta=max(data["gap"], data["pinf"], data["dinf"]), hv=data["hv_count"],
                                    iters=data["iters"], n=int(K["s"]), m=int(At.shape[1]))
    with open(f"{HERE}/oracle_outputs.json", "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps({k: v for k, v in out.items() if "log" not in k}, indent=1))


if __name__ == "__main__":
    main()
