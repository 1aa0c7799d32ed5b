"""Host-side mirror of the reference's public drivers over the C ABI.

    [X, obj, data] = ManiSDP_onlyunitdiag(C, options)          src/primal/ManiSDP_onlyunitdiag.m:6
    [X, obj, data] = ManiSDP_unitdiag(At, b, c, K, options)    src/primal/ManiSDP_unitdiag.m:7
    [X, obj, data] = ManiSDP_unittrace(At, b, c, K, options)   src/primal/ManiSDP_unittrace.m:7
    [X, obj, data] = ManiSDP(At, b, c, K, options)             src/primal/ManiSDP.m:6

Same names, argument meaning, option fields / defaults, console lines and `data` fields as the
MATLAB functions; matlab/*.m are the same few lines written against the MEX gateway.  Everything n-sized runs in
libmanisdp_b200.so on the GPU: the trust-region solve (trustregions + tCG + closures), the KKT residues, the
eigen step, the rank step and the escape update.  This file only sequences those calls and applies the scalar rules of
the outer loop (stopping test, sigma rule, slow-progress abort).  It never imports the oracle and has no CPU path.
"""
from __future__ import annotations

import time

import numpy as np

from . import _lib

DEFAULTS = {
    # ManiSDP_onlyunitdiag.m:8-17
    "onlyunitdiag": dict(p0=2, AL_maxiter=20, tol=1e-8, theta=1e-1, delta=8, alpha=0.5, tolgradnorm=1e-8,
                         TR_maxinner=100, TR_maxiter=40, line_search=0),
    # ManiSDP_unitdiag.m:10-26
    "unitdiag": dict(p0=2, AL_maxiter=300, gama=2, sigma0=1e-3, sigma_min=1e-2, sigma_max=1e7, tol=1e-8, theta=1e-3,
                     delta=8, alpha=0.1, tolgradnorm=1e-8, TR_maxinner=20, TR_maxiter=4, tau1=1, tau2=1,
                     line_search=0),
    # ManiSDP_unittrace.m:10-25
    "unittrace": dict(p0=1, AL_maxiter=1000, gama=2, sigma0=1e1, sigma_min=1e2, sigma_max=1e7, tol=1e-8, theta=1e-2,
                      delta=8, alpha=0.05, tolgradnorm=1e-8, TR_maxinner=40, TR_maxiter=3, tau1=1e-5, tau2=1e-4,
                      line_search=1),
    # ManiSDP.m:9-25
    "general": dict(p0=1, AL_maxiter=1000, gama=2, sigma0=1e-2, sigma_min=1e-1, sigma_max=1e7, tol=1e-8, theta=1e-2,
                    delta=8, alpha=0.1, tolgradnorm=1e-8, TR_maxinner=20, TR_maxiter=4, tau1=1e-2, tau2=1e-1,
                    line_search=1),
}
# largest n for which the dense outputs X = YY' and data.S are formed (the reference always forms them; at
# n = 1e6 that is 8 TB, so large problems return data['Y'] instead -- documented extension)
DENSE_OUTPUT_MAX_N = 6000


def _opts(kind, options):
    o = dict(DEFAULTS[kind])
    o.update(options or {})
    o.setdefault("seed", 0)
    o.setdefault("verbose", True)
    o.setdefault("use_graph", 1)
    o.setdefault("eig_tol", 0.0)  # 0: adaptive eigen-step accuracy keyed to options.tol (include/manisdp_b200.h: manisdp_kkt)
    o.setdefault("device", 0)
    return o


def _say(o, msg):
    if o["verbose"]:
        print(msg, flush=True)


def _init_point(h, o):
    Y0 = o.get("Y0")
    if Y0 is not None:
        h.set_Y(np.asarray(Y0, dtype=np.float64))
    else:
        h.rand_Y(int(o["p0"]), int(o["seed"]))  # trustregions.m:390-392 -> M.rand()


def ManiSDP_onlyunitdiag(C, options=None):
    """min <C,X> s.t. diag(X) = 1, X >= 0  (src/primal/ManiSDP_onlyunitdiag.m)."""
    import scipy.sparse as sp

    o = _opts("onlyunitdiag", options)
    C = sp.csc_matrix(C)
    n = C.shape[0]
    _say(o, "ManiSDP is starting...")
    _say(o, f"SDP size: n = {n}, m = {n}")
    data = dict(status=0, hv_count=0, tr_iters=0, fac_size=[], tr_seconds=0.0)
    t0 = time.perf_counter()
    # Multi-GPU (extension): options.world > 1 with options.rank / options.nccl_id runs the same loop on a
    # column-sharded handle -- the trust-region solve on p/world columns per GPU, the outer-loop steps (eigen step, rank
    # step, escape) redundantly and identically on every rank on the merged factor.  One process per GPU calls this.
    # options.devices = [d0, d1, ...] instead: ONE process drives all the listed GPUs (manisdp_group_*, csrc/group.cu) --
    # the mode the MATLAB gateway uses; the group splits / merges inside its tr_solve.
    world = int(o.get("world", 1))
    devices = o.get("devices")
    if devices is not None and len(devices) > 0:
        world = 1
        opener = _lib.GroupHandle(n, C, list(devices))
    else:
        opener = _lib.Handle("onlyunitdiag", n, C_csc=C, device=o["device"], rank=int(o.get("rank", 0)), world=world,
                             nccl_id=o.get("nccl_id"), layout="cols" if world > 1 else "rows")
    with opener as h:
        _init_point(h, o)
        data["setup_seconds"] = time.perf_counter() - t0
        staged = False
        dinf0 = None
        for it in range(1, int(o["AL_maxiter"]) + 1):
            data["fac_size"].append(h.p)
            if staged:
                h.line_search()  # :40-42
            if world > 1:
                h.col_split()
            info = h.tr_solve(o["TR_maxiter"], o["TR_maxinner"], o["tolgradnorm"], o["use_graph"])  # :43
            if world > 1:
                h.col_merge()
            data["hv_count"] += info.hv_count
            data["tr_iters"] += info.iters
            data["tr_seconds"] += info.seconds
            gradnorm = info.gradnorm
            t_k = time.perf_counter()
            k = h.kkt(int(o["delta"]), o["eig_tol"] if o["eig_tol"] > 0 else -float(o["tol"]), 0)  # :45-51
            data["kkt_seconds"] = data.get("kkt_seconds", 0.0) + time.perf_counter() - t_k
            data["eig_iters_total"] = data.get("eig_iters_total", 0) + int(k.eig_iters)
            data["eig_unconverged"] = data.get("eig_unconverged", 0) + (0 if k.eig_converged else 1)
            obj, dinf = k.obj, k.dinf
            p = h.p
            r, _ = h.rank_cut(o["theta"], apply=False)  # :52-54
            _say(o, f"Iter {it}, obj:{obj:0.8f}, dinf:{dinf:0.1e}, r:{r}, p:{p}, time:{time.perf_counter()-t0:0.2f}s")
            if dinf < o["tol"]:
                _say(o, "Optimality is reached!")
                break
            if it % 20 == 0:  # :61-69
                if it > 50 and dinf > dinf0:
                    data["status"] = 2
                    _say(o, "Slow progress!")
                    break
                dinf0 = dinf
            if it == int(o["AL_maxiter"]):
                break  # the reference still updates Y here, but nothing consumes it afterwards
            if r <= p - 1:
                h.rank_cut(o["theta"], apply=True)  # :70-73
            nne = max(min(k.nneg, int(o["delta"])), 1)  # :74
            staged = int(o["line_search"]) == 1
            h.escape(nne, o["alpha"], int(o["line_search"]))  # :75-84
        Y = h.get_Y()
        st = h.stats()
        data["launches"] = st.launches_total
        data["n_devices"] = len(devices) if devices else world
    z = None
    X = S = None
    if n <= DENSE_OUTPUT_MAX_N:
        X = Y @ Y.T
        z = np.asarray(C.multiply(X).sum(axis=0)).ravel()
        S = C.toarray() - np.diag(z)
    data.update(X=X, S=S, z=z, dinf=dinf, gradnorm=gradnorm, time=time.perf_counter() - t0, Y=Y, iters=it, obj=obj,
                lam_min=k.lam_min, lam_max=k.lam_max, eig_iters=k.eig_iters, eig_resid=k.eig_resid,
                eig_converged_last=int(k.eig_converged))
    if data["status"] == 0 and dinf > o["tol"]:
        data["status"] = 1
        _say(o, "Iteration maximum is reached!")
    _say(o, f"ManiSDP: optimum = {obj:0.8f}, time = {data['time']:0.2f}s")
    return X, obj, data


def _affine_driver(kind, At, b, c, K, options):
    import scipy.sparse as sp

    o = _opts(kind, options)
    n = int(K["s"] if isinstance(K, dict) else K)
    At = sp.csc_matrix(At)
    bd = np.asarray(b.todense()).ravel() if sp.issparse(b) else np.asarray(b, dtype=np.float64).ravel()
    m = At.shape[1]
    _say(o, "ManiSDP is starting...")
    _say(o, f"SDP size: n = {n}, m = {m}")
    sigma = float(o["sigma0"])
    gama = float(o["gama"])
    data = dict(status=0, hv_count=0, tr_iters=0, fac_size=[], tr_seconds=0.0)
    check_every, check_after = (50, 100) if kind == "unitdiag" else (20, 50)
    t0 = time.perf_counter()
    gap0 = pinf0 = dinf0 = None
    phase = dict(create=0.0, line_search=0.0, tr_solve=0.0, kkt=0.0, rank=0.0, escape=0.0)  # wall seconds per phase

    def timed(name, fn, *a):
        t1 = time.perf_counter()
        r_ = fn(*a)
        phase[name] += time.perf_counter() - t1
        return r_

    with _lib.Handle(kind, n, At=At, b=bd, c=c, device=o["device"], force_mode=int(o.get("force_mode", 0))) as h:
        h.set_dual(np.zeros(m), sigma)
        _init_point(h, o)
        phase["create"] = time.perf_counter() - t0
        staged = False
        for it in range(1, int(o["AL_maxiter"]) + 1):
            data["fac_size"].append(h.p)
            if staged:
                timed("line_search", h.line_search)
            info = timed("tr_solve", h.tr_solve, o["TR_maxiter"], o["TR_maxinner"], o["tolgradnorm"], o["use_graph"])
            data["hv_count"] += info.hv_count
            data["tr_iters"] += info.iters
            data["tr_seconds"] += info.seconds
            gradnorm = info.gradnorm
            k = timed("kkt", h.kkt, int(o["delta"]), o["eig_tol"] if o["eig_tol"] > 0 else -float(o["tol"]), 1)  # residues, y <- y - sigma*Axb, eig(S)
            obj, gap, pinf, dinf = k.obj, k.gap, k.pinf, k.dinf
            # Ritz values bound lambda_min from above: an eigen step that missed its residual test (after the engine's
            # own two warm restarts) is counted, so a reported dinf can be audited (data['eig_unconverged'])
            data["eig_unconverged"] = data.get("eig_unconverged", 0) + (0 if k.eig_converged else 1)
            data["eig_iters_total"] = data.get("eig_iters_total", 0) + int(k.eig_iters)
            p = h.p
            r, _ = timed("rank", h.rank_cut, o["theta"], False)
            _say(o, f"Iter {it}, obj:{obj:0.8f}, gap:{gap:0.1e}, pinf:{pinf:0.1e}, dinf:{dinf:0.1e}, "
                    f"gradnorm:{gradnorm:0.1e}, r:{r}, p:{p}, sigma:{sigma:0.3f}, time:{time.perf_counter()-t0:0.2f}s")
            eta = max(gap, pinf, dinf)
            if eta < o["tol"]:
                _say(o, "Optimality is reached!")
                break
            if it % check_every == 0:
                if it > check_after and gap > gap0 and pinf > pinf0 and dinf > dinf0:
                    data["status"] = 2
                    _say(o, "Slow progress!")
                    break
                gap0, pinf0, dinf0 = gap, pinf, dinf
            if it == int(o["AL_maxiter"]):
                break
            if r <= p - 1:
                timed("rank", h.rank_cut, o["theta"], True)
            nne = min(k.nneg, int(o["delta"]))
            if kind == "unitdiag":
                nne = max(nne, 1)  # ManiSDP_unitdiag.m:97 (the other two have no lower bound)
            staged = int(o["line_search"]) == 1
            timed("escape", h.escape, nne, o["alpha"], int(o["line_search"]))
            if pinf < o["tau1"] * gradnorm:  # ManiSDP_unitdiag.m:108-112
                sigma = max(sigma / gama, o["sigma_min"])
            elif pinf > o["tau2"] * gradnorm:
                sigma = min(sigma * gama, o["sigma_max"])
            h.set_sigma(sigma)
        Y = h.get_Y()
        y, _ = h.get_dual()
        st = h.stats()
        data["launches"] = st.launches_total
        data["s_mode"], data["a_mode"] = st.s_mode, st.a_mode
        data["phase_seconds"] = phase
    X = S = z = None
    if n <= DENSE_OUTPUT_MAX_N:
        X = Y @ Y.T
        cd = np.asarray(c.todense()).ravel() if sp.issparse(c) else np.asarray(c, dtype=np.float64).ravel()
        eS = (cd - At @ y).reshape(n, n, order="F")
        if kind == "unitdiag":
            z = np.sum(X * eS, axis=0)
            S = eS - np.diag(z)
        elif kind == "unittrace":
            z = float(np.sum(X * eS))
            S = eS - z * np.eye(n)
        else:
            S = eS
    data.update(X=X, y=y, S=S, z=z, gap=gap, pinf=pinf, dinf=dinf, gradnorm=gradnorm, time=time.perf_counter() - t0,
                Y=Y, iters=it, obj=obj, sigma=sigma, lam_min=k.lam_min, lam_max=k.lam_max, eig_iters=k.eig_iters,
                eig_resid=k.eig_resid, eig_converged_last=int(k.eig_converged))
    if data["status"] == 0 and eta > o["tol"]:
        data["status"] = 1
        _say(o, "Iteration maximum is reached!")
    _say(o, f"ManiSDP: optimum = {obj:0.8f}, time = {data['time']:0.2f}s")
    return X, obj, data


def ManiSDP_unitdiag(At, b, c, K, options=None):
    """src/primal/ManiSDP_unitdiag.m:7"""
    return _affine_driver("unitdiag", At, b, c, K, options)


def ManiSDP_unittrace(At, b, c, K, options=None):
    """src/primal/ManiSDP_unittrace.m:7"""
    return _affine_driver("unittrace", At, b, c, K, options)


def ManiSDP(At, b, c, K, options=None):
    """src/primal/ManiSDP.m:6"""
    return _affine_driver("general", At, b, c, K, options)



# ManiSDP_multiblock.m:10-27
MB_DEFAULTS = dict(min_facsize=2, p0=None, AL_maxiter=1000, gama=2, sigma0=1e-1, sigma_min=1e-2, sigma_max=1e7,
                   tol=1e-8, theta=1e-2, delta=8, alpha=0.1, tolgradnorm=1e-8, TR_maxinner=20, TR_maxiter=4,
                   tau1=1e1, tau2=1e1, line_search=0)


def ManiSDP_multiblock(At, b, c, K, options=None):
    """min <C,X> s.t. A(X) = b, X in S_+^{n_1 x ... x n_t}, diag(X_i) = 1 for i <= K.nob
    (src/primal/ManiSDP_multiblock.m:7).  K['s']: block orders, K['nob']: number of leading unit-diagonal blocks.
    options['Y0'] (extension): list of (n_i, p_i) arrays.  Returns X and data['S'] as lists of blocks."""
    import scipy.sparse as sp

    o = dict(MB_DEFAULTS)
    o.update(options or {})
    for k_, v_ in dict(seed=0, verbose=True, use_graph=1, device=0).items():
        o.setdefault(k_, v_)
    nset = [int(v) for v in np.atleast_1d(K["s"])]
    nb = len(nset)
    nob = int(K.get("nob", 0))
    N = int(sum(nset))
    At = sp.csc_matrix(At)
    bd = np.asarray(b.todense()).ravel() if sp.issparse(b) else np.asarray(b, dtype=np.float64).ravel()
    m = At.shape[1]
    _say(o, "ManiSDP is starting...")
    _say(o, f"SDP size: n = {max(nset)}, m = {m}")
    p0 = o["p0"] if o["p0"] is not None else [1] * nb  # :12
    p = list(nset)  # :33-39
    for i in range(nb):
        if nset[i] >= o["min_facsize"]:
            p[i] = int(p0[i])
    sigma, gama = float(o["sigma0"]), float(o["gama"])
    data = dict(status=0, hv_count=0, tr_iters=0, fac_size=[], tr_seconds=0.0)
    phase = dict(create=0.0, line_search=0.0, tr_solve=0.0, kkt=0.0, update=0.0)
    t0 = time.perf_counter()
    gap0 = pinf0 = dinf0 = None

    def timed(name, fn, *a):
        t1 = time.perf_counter()
        r_ = fn(*a)
        phase[name] += time.perf_counter() - t1
        return r_

    with _lib.Handle("multiblock", N, At=At, b=bd, c=c, device=o["device"], block_sizes=nset, nob=nob) as h:
        h.set_dual(np.zeros(m), sigma)
        if o.get("Y0") is not None:
            h.mb_set_Y([np.asarray(B, dtype=np.float64) for B in o["Y0"]])
        else:
            h.mb_rand_Y(p, int(o["seed"]))  # trustregions.m:390-392 -> M.rand() (randc.cpp)
        phase["create"] = time.perf_counter() - t0
        staged = False
        for it in range(1, int(o["AL_maxiter"]) + 1):
            p = h.mb_widths()
            data["fac_size"].append(list(p))
            if staged:
                timed("line_search", h.line_search)  # :62-64
            info = timed("tr_solve", h.tr_solve, o["TR_maxiter"], o["TR_maxinner"], o["tolgradnorm"], o["use_graph"])
            data["hv_count"] += info.hv_count
            data["tr_iters"] += info.iters
            data["tr_seconds"] += info.seconds
            gradnorm = info.gradnorm
            k, dinfs, nneg = timed("kkt", h.mb_kkt, 1)  # :66-97
            obj, gap, pinf, dinf = k.obj, k.gap, k.pinf, k.dinf
            _say(o, f"Iter {it}, obj:{obj:0.8f}, gap:{gap:0.1e}, pinf:{pinf:0.1e}, dinf:{dinf:0.1e}, "
                    f"gradnorm:{gradnorm:0.1e}, p_max:{max(p)}, sigma:{sigma:0.3f}, time:{time.perf_counter()-t0:0.2f}s")
            eta = max(gap, pinf, dinf)
            if eta < o["tol"]:
                _say(o, "Optimality is reached!")
                break
            if it % 50 == 0:  # :103-113
                if it > 100 and gap > gap0 and pinf > pinf0 and dinf > dinf0:
                    data["status"] = 2
                    _say(o, "Slow progress!")
                    break
                gap0, pinf0, dinf0 = gap, pinf, dinf
            if it == int(o["AL_maxiter"]):
                break
            timed("update", h.mb_update, o["theta"], o["delta"], o["alpha"], int(o["line_search"]),
                  int(o["min_facsize"]))  # :114-153
            staged = int(o["line_search"]) == 1
            if pinf < o["tau1"] * gradnorm:  # :154-158
                sigma = max(sigma / gama, o["sigma_min"])
            elif pinf > o["tau2"] * gradnorm:
                sigma = min(sigma * gama, o["sigma_max"])
            h.set_sigma(sigma)
        Y = h.mb_get_Y()
        y, _ = h.get_dual()
        st = h.stats()
        data["launches"] = st.launches_total
        data["phase_seconds"] = phase
    X = [Yi @ Yi.T for Yi in Y]
    cd = np.asarray(c.todense()).ravel() if sp.issparse(c) else np.asarray(c, dtype=np.float64).ravel()
    cy = cd - At @ y
    S, o2 = [], 0
    for i, ni in enumerate(nset):
        Si = cy[o2:o2 + ni * ni].reshape(ni, ni, order="F")
        if i < nob:
            Si = Si - np.diag(np.sum(X[i] * Si, axis=0))
        S.append(Si)
        o2 += ni * ni
    data.update(X=X, y=y, S=S, gap=gap, pinf=pinf, dinf=dinf, gradnorm=gradnorm, time=time.perf_counter() - t0, Y=Y,
                iters=it, obj=obj, sigma=sigma, dinfs=dinfs)
    if data["status"] == 0 and eta > o["tol"]:
        data["status"] = 1
        _say(o, "Iteration maximum is reached!")
    _say(o, f"ManiSDP: optimum = {obj:0.8f}, time = {data['time']:0.2f}s")
    return X, obj, data


# ManiDSDP_unitdiag.m:10-26
DUAL_DEFAULTS = dict(p0=None, ADMM_maxiter=300, gama=2, sigma0=1e-3, sigma_min=1e-3, sigma_max=1e7, tol=1e-8, theta=1e-3,
                     delta=8, alpha=0.1, tolgradnorm=1e-8, TR_maxinner=20, TR_maxiter=4, tau1=1e1, tau2=1e2,
                     line_search=0)


def ManiDSDP_unitdiag(A, b, c, K, options=None):
    """sup <C,X> + <c_f,w> s.t. A(X) + B(w) = b, X >= 0, with a unit-diagonal dual slack -- the dual approach
    (Riemannian ADMM) of src/dual/ManiDSDP_unitdiag.m:8.  A: (m, K.f + n*n), c: (K.f + n*n,), K = {'f': .., 's': n},
    options['dAAt'] = diag(A*A') when known (bqpsos returns it)."""
    import math

    import scipy.sparse as sp

    o = dict(DUAL_DEFAULTS)
    o.update(options or {})
    for k_, v_ in dict(seed=0, verbose=True, use_graph=1, device=0, eig_tol=0.0).items():
        o.setdefault(k_, v_)
    n = int(K["s"])
    nf = int(K.get("f", 0))
    A = sp.csc_matrix(A)
    bd = np.asarray(b, dtype=np.float64).ravel()
    cd = np.asarray(c.todense()).ravel() if sp.issparse(c) else np.asarray(c, dtype=np.float64).ravel()
    m = A.shape[0]
    _say(o, "ManiSDP is starting...")
    _say(o, f"SDP size: n = {n}, m = {m}")
    if o["p0"] is None:
        o["p0"] = int(math.ceil(math.log(m)))  # :11
    B, Ap = A[:, :nf], A[:, nf:]  # :34-37
    cf, cp = cd[:nf], cd[nf:]
    sigma, gama = float(o["sigma0"]), float(o["gama"])
    data = dict(status=0, hv_count=0, tr_iters=0, fac_size=[], seta=[], tr_seconds=0.0)
    phase = dict(create=0.0, line_search=0.0, tr_solve=0.0, kkt=0.0, rank=0.0, escape=0.0)
    t0 = time.perf_counter()
    gap0 = pinf0 = dinf0 = None

    def timed(name, fn, *a):
        t1 = time.perf_counter()
        r_ = fn(*a)
        phase[name] += time.perf_counter() - t1
        return r_

    with _lib.Handle("dual_unitdiag", n, At=Ap.T.tocsc(), b=bd, c=cp, device=o["device"], dAAt=o.get("dAAt"), B=B,
                     cf=cf, force_mode=int(o.get("force_mode", 0))) as h:
        h.set_sigma(sigma)
        _init_point(h, o)
        phase["create"] = time.perf_counter() - t0
        staged = False
        for it in range(1, int(o["ADMM_maxiter"]) + 1):
            data["fac_size"].append(h.p)
            if staged:
                timed("line_search", h.line_search)  # :66-68
            info = timed("tr_solve", h.tr_solve, o["TR_maxiter"], o["TR_maxinner"], o["tolgradnorm"], o["use_graph"])
            data["hv_count"] += info.hv_count
            data["tr_iters"] += info.iters
            data["tr_seconds"] += info.seconds
            gradnorm = info.gradnorm
            # ADMM step + residues + eig(X)  (:71-88)
            k = timed("kkt", h.kkt, int(o["delta"]), o["eig_tol"] if o["eig_tol"] > 0 else -float(o["tol"]), 1)
            obj, gap, pinf, dinf = k.obj, k.gap, k.pinf, k.dinf
            data["eig_unconverged"] = data.get("eig_unconverged", 0) + (0 if k.eig_converged else 1)
            p = h.p
            r, _ = timed("rank", h.rank_cut, o["theta"], False)  # :89-91
            _say(o, f"Iter {it}, obj:{obj:0.8f}, gap:{gap:0.1e}, pinf:{pinf:0.1e}, dinf:{dinf:0.1e}, "
                    f"gradnorm:{gradnorm:0.1e}, r:{r}, p:{p}, sigma:{sigma:0.3f}, time:{time.perf_counter()-t0:0.2f}s")
            eta = max(gap, pinf, dinf)
            data["seta"].append(eta)
            if eta < o["tol"]:
                _say(o, "Optimality is reached!")
                break
            if it % 50 == 0:
                if it > 100 and gap > gap0 and pinf > pinf0 and dinf > dinf0:
                    data["status"] = 2
                    _say(o, "Slow progress!")
                    break
                gap0, pinf0, dinf0 = gap, pinf, dinf
            if it == int(o["ADMM_maxiter"]):
                break
            if r <= p - 1:  # :112-115
                timed("rank", h.rank_cut, o["theta"], True)
            nne = max(min(k.nneg, int(o["delta"])), 1)  # :116
            staged = int(o["line_search"]) == 1
            timed("escape", h.escape, nne, o["alpha"], int(o["line_search"]))
            if pinf < o["tau1"] * gradnorm:  # :128-132
                sigma = max(sigma / gama, o["sigma_min"])
            elif pinf > o["tau2"] * gradnorm:
                sigma = min(sigma * gama, o["sigma_max"])
            h.set_sigma(sigma)
        Y = h.get_Y()
        y, _ = h.get_dual()
        x, w = h.dual_state()
        st = h.stats()
        data["launches"] = st.launches_total
        data["s_mode"], data["a_mode"] = st.s_mode, st.a_mode
        data["phase_seconds"] = phase
    X = S = None
    if n <= DENSE_OUTPUT_MAX_N:
        S = Y @ Y.T
        dA = np.asarray(o["dAAt"], dtype=np.float64).ravel() if o.get("dAAt") is not None else \
            np.asarray(Ap.multiply(Ap).sum(axis=1)).ravel()
        eX = (x + Ap.T @ (bd / dA)).reshape(n, n, order="F")
        X = eX - np.diag(np.sum(S * eX, axis=0))
    data.update(X=X, y=y, S=S, w=w, x=x, gap=gap, pinf=pinf, dinf=dinf, gradnorm=gradnorm, time=time.perf_counter() - t0,
                Y=Y, iters=it, obj=obj, sigma=sigma, lam_min=k.lam_min, lam_max=k.lam_max)
    if data["status"] == 0 and eta > o["tol"]:
        data["status"] = 1
        _say(o, "Iteration maximum is reached!")
    _say(o, f"ManiDSDP: optimum = {obj:0.8f}, time = {data['time']:0.2f}s")
    return X, obj, data
