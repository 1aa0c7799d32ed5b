"""ORACLE (test infrastructure, not product code) -- the four ManiSDP primal drivers.

CPU restatement (NumPy/SciPy FP64, dense X / dense eig exactly as the reference does) of

  * src/primal/ManiSDP_onlyunitdiag.m   (closures :117-130, outer loop :38-84, line search :101-115)
  * src/primal/ManiSDP_unitdiag.m       (closures :152-171, outer loop :51-113, line search :131-150)
  * src/primal/ManiSDP_unittrace.m      (closures :156-177, outer loop :52-117)
  * src/primal/ManiSDP.m                (closures :149-165, outer loop :52-113)
  * src/primal/ManiSDP_multiblock.m     (closures :203-247, outer loop :60-160, line search :180-201)
  * src/dual/ManiDSDP_unitdiag.m        (closures :171-192, outer loop :63-141, line search :158-169)

driving oracle/manopt_rtr.py in place of Manopt's trustregions.  It is the parity yardstick for the
CUDA engine and the CPU baseline timed by bench.py; nothing in the product path imports it.

PARITY PIN: the reference ships no expected optima and cannot run here (no
MATLAB/Octave).  The oracle is pinned by tests/test_oracle_kats.py against independent known
answers: SDPLIB optimal values listed in data/sdplib/README (maxG11 629.1648, ...), a brute-force
BQP minimum, DIMACS theta values of Hamming graphs, and the survey's recorded optima
(BASELINE.md section 2).

LAYOUT: the factor is always (n, p) with one ROW per vertex (see oracle/manopt_rtr.py).  `At` is
scipy CSC (n*n, m) over the column-major vec of X.

Extensions over the reference (documented, needed for deterministic parity tests):
  * options['Y0'] accepted by all four drivers (reference: only unittrace / ManiSDP, :36-40);
  * options['seed'] selects the NumPy generator used where the reference calls randn;
  * options['stale_eG'] (onlyunitdiag only, default True = reference behaviour): the reference's
    hess closure reads the shared variable eG that cost() overwrites at a REJECTED proposal
    (ManiSDP_onlyunitdiag.m:118-119,129).  False keeps eG of the accepted point.
  * data['hv_count'], data['tr_iters'], data['iters'] counters (the reference has none).
"""
from __future__ import annotations

import math
import time

import numpy as np
import scipy.sparse as sp

from .manopt_rtr import Cells, Euclid, MultiBlock, ObliqueT, Sphere, trustregions

# option defaults, SURVEY.md Appendix B (source lines cited there)
DEFAULTS = {
    "onlyunitdiag": dict(p0=2, AL_maxiter=20, tol=1e-8, theta=1e-1, delta=8, alpha=0.5,
                         tolgradnorm=1e-8, TR_maxinner=100, TR_maxiter=40, line_search=0),
    "unitdiag": dict(p0=2, AL_maxiter=300, gama=2, sigma0=1e-3, sigma_min=1e-2, sigma_max=1e7,
                     tol=1e-8, theta=1e-3, delta=8, alpha=0.1, tolgradnorm=1e-8, TR_maxinner=20,
                     TR_maxiter=4, tau1=1, tau2=1, line_search=0),
    "unittrace": dict(p0=1, AL_maxiter=1000, gama=2, sigma0=1e1, sigma_min=1e2, sigma_max=1e7,
                      tol=1e-8, theta=1e-2, delta=8, alpha=0.05, tolgradnorm=1e-8, TR_maxinner=40,
                      TR_maxiter=3, tau1=1e-5, tau2=1e-4, line_search=1),
    "general": dict(p0=1, AL_maxiter=1000, gama=2, sigma0=1e-2, sigma_min=1e-1, sigma_max=1e7,
                    tol=1e-8, theta=1e-2, delta=8, alpha=0.1, tolgradnorm=1e-8, TR_maxinner=20,
                    TR_maxiter=4, tau1=1e-2, tau2=1e-1, line_search=1),
}


def _opts(kind, options):
    o = dict(DEFAULTS[kind])
    o.update(options or {})
    o.setdefault("seed", 0)
    o.setdefault("verbose", False)
    return o


def _mat(v, n):
    """reshape(v, n, n) of a column-major vec."""
    return np.asarray(v).reshape(n, n, order="F")


def _vec(Xm):
    return np.asarray(Xm).reshape(-1, order="F")


# --------------------------------------------------------------------------------------------
# closures
# --------------------------------------------------------------------------------------------
class OnlyUnitDiagProblem:
    """ManiSDP_onlyunitdiag.m:117-130 (f = 0.5 <C, Y Y'>)."""

    def __init__(self, C, p, stale_eG=True):
        self.C = sp.csr_matrix(C)
        self.M = ObliqueT(self.C.shape[0], p)
        self.stale = stale_eG
        self.YC = self.eG = None  # state of the last cost() call (the reference's shared variables)
        self._kept = None

    def cost(self, Y):
        self._prev = (self.YC, self.eG)
        self.YC = self.C @ Y  # :118  (C symmetric: (Y_ref*C)' = C*Y)
        self.eG = np.sum(self.YC * Y, axis=1, keepdims=True)  # :119
        return 0.5 * float(self.eG.sum())  # :120

    def accept(self, ok):
        if not ok and not self.stale:
            self.YC, self.eG = self._prev  # keep the accepted point's eG (correct maths)

    def grad(self, Y):
        return self.YC - Y * self.eG  # :124

    def hess(self, Y, U):
        eH = self.C @ U  # :128
        return eH - Y * np.sum(Y * eH, axis=1, keepdims=True) - U * self.eG  # :129


class AffineProblem:
    """Closures of ManiSDP_unitdiag.m:152-171 (kind 'unitdiag'), ManiSDP_unittrace.m:156-177
    ('unittrace') and ManiSDP.m:149-165 ('general') on f(Y) = c'x + sigma/2 |A x - b - y/sigma|^2,
    x = vec(Y Y')."""

    def __init__(self, kind, At, b, c, n, p, y, sigma):
        self.kind, self.n = kind, n
        self.At = At
        self.A = At.T.tocsr()
        self.b, self.c, self.y, self.sigma = b, c, y, sigma
        self.M = {"unitdiag": ObliqueT, "unittrace": Sphere, "general": Euclid}[kind](n, p)
        self.Axb = None

    # cost at an arbitrary point without touching the closure state (the reference's `co`)
    def co(self, Y):
        x = _vec(Y @ Y.T)
        Axb = self.A @ x - self.b - self.y / self.sigma
        return float(self.c @ x + self.sigma / 2 * (Axb @ Axb))

    def cost(self, Y):
        x = _vec(Y @ Y.T)
        self.Axb = self.A @ x - self.b - self.y / self.sigma
        self._Ycost = Y
        return float(self.c @ x + 0.5 * self.sigma * (self.Axb @ self.Axb))

    def accept(self, ok):
        pass  # every closure that reads shared state is only called after an accepted cost()

    def grad(self, Y):
        n = self.n
        self.eS = _mat(self.c + self.sigma * (self.At @ self.Axb), n)
        if self.kind == "unitdiag":  # ManiSDP_unitdiag.m:159-164
            eG = 2 * (self.eS @ Y)
            self.YeG = np.sum(Y * eG, axis=1, keepdims=True)
            return eG - Y * self.YeG
        if self.kind == "unittrace":  # ManiSDP_unittrace.m:161-164 (built inside cost there)
            self.z = float(np.sum((self.eS @ Y) * Y))
            return 2 * (self.eS @ Y) - 2 * self.z * Y
        return 2 * (self.eS @ Y)  # ManiSDP.m:156-159

    def hess(self, Y, U):
        n, s = self.n, self.sigma
        if self.kind == "unitdiag":  # ManiSDP_unitdiag.m:166-171 : YU(i,j) = <Y_i, U_j>
            YU = Y @ U.T
        else:  # ManiSDP_unittrace.m:172 / ManiSDP.m:162 : YU = U*Y'
            YU = U @ Y.T
        AyU = _mat(self.At @ (self.A @ _vec(YU)), n)
        H = 2 * (self.eS @ U) + 4 * s * (AyU @ Y)
        if self.kind == "unitdiag":
            return H - Y * np.sum(Y * H, axis=1, keepdims=True) - U * self.YeG
        if self.kind == "unittrace":  # :174-176
            return H - float(np.vdot(H, Y)) * Y - 2 * self.z * U
        return H


# --------------------------------------------------------------------------------------------
# shared pieces of the outer loops
# --------------------------------------------------------------------------------------------
def _normalize(kind, Y):
    if kind in ("onlyunitdiag", "unitdiag"):
        return Y / np.sqrt(np.sum(Y * Y, axis=1, keepdims=True))
    if kind == "unittrace":
        return Y / np.linalg.norm(Y)
    return Y


def _line_search(kind, co, Y, U):
    """ManiSDP_unitdiag.m:138-150 (identical in the other three drivers up to the normalisation)."""
    alpha = 1.0
    cost0 = co(Y)
    i = 1
    nY = _normalize(kind, Y + alpha * U)
    while i <= 15 and co(nY) - cost0 > -1e-3:
        alpha *= 0.8
        nY = _normalize(kind, Y + alpha * U)
        i += 1
    return nY


def _rank_cut(Y, theta):
    """svd(Y) rank estimate and truncation (ManiSDP_unitdiag.m:72-74,93-96 in row layout):
    returns (r, Y_r) with Y_r = U_r diag(e_r) (n x r)."""
    Us, e, _ = np.linalg.svd(Y, full_matrices=False)
    r = int(np.sum(e >= theta * e[0]))
    return r, Us[:, :r] * e[:r]


def _escape(kind, o, Y, vS, nne):
    """append nne escape directions (ManiSDP_unitdiag.m:97-107 and siblings).  Returns (Y, U)."""
    n, p = Y.shape
    V = vS[:, :nne]
    U = None
    if o["line_search"] == 1:
        U = np.hstack([np.zeros((n, p)), V])
        Y = np.hstack([Y, np.zeros((n, nne))])
    else:
        Y = _normalize(kind, np.hstack([Y, o["alpha"] * V]))
    return Y, U


# --------------------------------------------------------------------------------------------
# drivers
# --------------------------------------------------------------------------------------------
def ManiSDP_onlyunitdiag(C, options=None):
    """[X, obj, data] of src/primal/ManiSDP_onlyunitdiag.m:6."""
    o = _opts("onlyunitdiag", options)
    C = sp.csr_matrix(C)
    n = C.shape[0]
    rng = np.random.default_rng(o["seed"])
    p = o["p0"]
    Y = o.get("Y0")
    if Y is not None:
        Y = np.array(Y, dtype=np.float64)
        p = Y.shape[1]
    U = None
    data = dict(status=0, hv_count=0, tr_iters=0, fac_size=[])
    t0 = time.perf_counter()
    dinf0 = None
    Cd = None
    for it in range(1, o["AL_maxiter"] + 1):
        data["fac_size"].append(p)
        prob = OnlyUnitDiagProblem(C, p, o.get("stale_eG", True))
        if Y is None:
            Y = prob.M.rand(rng)  # trustregions.m:390-392
        if U is not None:
            Y = _line_search("onlyunitdiag", lambda Z: float(np.sum((C @ Z) * Z)), Y, U)  # :101-115
        res = trustregions(prob, Y, maxiter=o["TR_maxiter"], maxinner=o["TR_maxinner"],
                           tolgradnorm=o["tolgradnorm"])
        Y = res.x
        data["hv_count"] += res.hv_count
        data["tr_iters"] += len(res.info) - 1
        gradnorm = res.info[-1].gradnorm
        X = Y @ Y.T  # :45
        if Cd is None:
            Cd = C.toarray()
        z = np.sum(Cd * X, axis=0)  # :46
        obj = float(z.sum())
        S = Cd - np.diag(z)
        dS, vS = np.linalg.eigh(S)  # :50
        dinf = max(0.0, -dS[0]) / (1 + dS[-1])
        r, Yr = _rank_cut(Y, o["theta"])
        if o["verbose"]:
            print(f"Iter {it}, obj:{obj:0.8f}, dinf:{dinf:0.1e}, r:{r}, p:{p}, time:{time.perf_counter()-t0:0.2f}s")
        if dinf < o["tol"]:
            break
        if it % 20 == 0:  # :61-69
            if it > 50 and dinf > dinf0:
                data["status"] = 2
                break
            dinf0 = dinf
        if r <= p - 1:
            Y, p = Yr, r
        nne = max(min(int(np.sum(dS < 0)), o["delta"]), 1)  # :74
        Y, U = _escape("onlyunitdiag", o, Y, vS, nne)
        p += nne
    data.update(X=X, S=S, z=z, dinf=dinf, gradnorm=gradnorm, time=time.perf_counter() - t0, Y=Y,
                iters=it, obj=obj)
    if data["status"] == 0 and dinf > o["tol"]:
        data["status"] = 1
    return X, obj, data


def _affine_driver(kind, At, b, c, K, options):
    o = _opts(kind, options)
    n = int(K["s"] if isinstance(K, dict) else K)
    At = sp.csc_matrix(At)
    b = np.asarray(b.todense()).ravel() if sp.issparse(b) else np.asarray(b, dtype=np.float64).ravel()
    c = np.asarray(c.todense()).ravel() if sp.issparse(c) else np.asarray(c, dtype=np.float64).ravel()
    A = At.T.tocsr()
    rng = np.random.default_rng(o["seed"])
    p = o["p0"]
    sigma = o["sigma0"]
    gama = o["gama"]
    y = np.zeros(len(b))
    normb = 1 + np.linalg.norm(b)
    Y = o.get("Y0")
    if Y is not None:
        Y = np.array(Y, dtype=np.float64)
        p = Y.shape[1]
    U = None
    data = dict(status=0, hv_count=0, tr_iters=0, fac_size=[])
    check_every, check_after = (50, 100) if kind == "unitdiag" else (20, 50)
    t0 = time.perf_counter()
    gap0 = pinf0 = dinf0 = None
    for it in range(1, o["AL_maxiter"] + 1):
        data["fac_size"].append(p)
        prob = AffineProblem(kind, At, b, c, n, p, y, sigma)
        if Y is None:
            Y = prob.M.rand(rng)
        if U is not None:
            Y = _line_search(kind, prob.co, Y, U)
        res = trustregions(prob, Y, maxiter=o["TR_maxiter"], maxinner=o["TR_maxinner"],
                           tolgradnorm=o["tolgradnorm"])
        Y = res.x
        data["hv_count"] += res.hv_count
        data["tr_iters"] += len(res.info) - 1
        gradnorm = res.info[-1].gradnorm
        X = Y @ Y.T
        x = _vec(X)
        obj = float(c @ x)
        Axb = A @ x - b
        pinf = float(np.linalg.norm(Axb)) / normb
        y = y - sigma * Axb
        eS = _mat(c - At @ y, n)
        if kind == "unitdiag":  # ManiSDP_unitdiag.m:65-70
            z = np.sum(X * eS, axis=0)
            S = eS - np.diag(z)
            by = float(b @ y + z.sum())
        elif kind == "unittrace":  # ManiSDP_unittrace.m:65-70
            z = float(np.sum(eS * X))
            S = eS - z * np.eye(n)
            by = float(b @ y + z)
        else:  # ManiSDP.m:64-68
            z = None
            S = eS
            by = float(b @ y)
        dS, vS = np.linalg.eigh(S)
        dinf = max(0.0, -dS[0]) / (1 + dS[-1])
        gap = abs(obj - by) / (abs(by) + abs(obj) + 1)
        r, Yr = _rank_cut(Y, o["theta"])
        if o["verbose"]:
            print(f"Iter {it}, obj:{obj:0.8f}, gap:{gap:0.1e}, pinf:{pinf:0.1e}, dinf:{dinf:0.1e}, "
                  f"gradnorm:{gradnorm:0.1e}, r:{r}, p:{p}, sigma:{sigma:0.3f}, "
                  f"time:{time.perf_counter()-t0:0.2f}s")
        eta = max(gap, pinf, dinf)
        if eta < o["tol"]:
            break
        if it % check_every == 0:
            if it > check_after and gap > gap0 and pinf > pinf0 and dinf > dinf0:
                data["status"] = 2
                break
            gap0, pinf0, dinf0 = gap, pinf, dinf
        if r <= p - 1:
            Y, p = Yr, r
        nneg = int(np.sum(dS < 0))
        nne = max(min(nneg, o["delta"]), 1) if kind == "unitdiag" else min(nneg, o["delta"])
        Y, U = _escape(kind, o, Y, vS, nne)
        p += nne
        if pinf < o["tau1"] * gradnorm:
            sigma = max(sigma / gama, o["sigma_min"])
        elif pinf > o["tau2"] * gradnorm:
            sigma = min(sigma * gama, o["sigma_max"])
    data.update(X=X, y=y, S=S, z=z, gap=gap, pinf=pinf, dinf=dinf, gradnorm=gradnorm,
                time=time.perf_counter() - t0, Y=Y, iters=it, obj=obj, sigma=sigma)
    if data["status"] == 0 and eta > o["tol"]:
        data["status"] = 1
    return X, obj, data


def ManiSDP_unitdiag(At, b, c, K, options=None):
    """[X, obj, data] of src/primal/ManiSDP_unitdiag.m:7."""
    return _affine_driver("unitdiag", At, b, c, K, options)


def ManiSDP_unittrace(At, b, c, K, options=None):
    """[X, obj, data] of src/primal/ManiSDP_unittrace.m:7."""
    return _affine_driver("unittrace", At, b, c, K, options)


def ManiSDP(At, b, c, K, options=None):
    """[X, obj, data] of src/primal/ManiSDP.m:6."""
    return _affine_driver("general", At, b, c, K, options)


# --------------------------------------------------------------------------------------------
# multi-block driver
# --------------------------------------------------------------------------------------------
MB_DEFAULTS = dict(min_facsize=2, p0=None, AL_maxiter=1000, gama=2, sigma0=1e-1, sigma_min=1e-2,
                   sigma_max=1e7, tol=1e-8, theta=1e-2, delta=8, alpha=0.1, tolgradnorm=1e-8,
                   TR_maxinner=20, TR_maxiter=4, tau1=1e1, tau2=1e1, line_search=0)  # :10-27


class MultiblockProblem:
    """Closures of ManiSDP_multiblock.m:203-247 on f(Y) = c'x + sigma/2 |A x - b - y/sigma|^2 with
    x = [vec(Y_1 Y_1'); ...; vec(Y_t Y_t')]  (block i: (n_i, p_i), row layout)."""

    def __init__(self, At, b, c, nset, pset, nob, y, sigma):
        self.At, self.A = At, At.T.tocsr()
        self.b, self.c, self.y, self.sigma = b, c, y, sigma
        self.n, self.nob = [int(v) for v in nset], int(nob)
        self.off = np.concatenate([[0], np.cumsum([v * v for v in self.n])]).astype(np.int64)
        self.M = MultiBlock(pset, nset, nob)

    def xvec(self, Y):  # :205-209
        return np.concatenate([_vec(Yi @ Yi.T) for Yi in Y])

    def blocks(self, v):
        return [_mat(v[self.off[i]:self.off[i + 1]], ni) for i, ni in enumerate(self.n)]

    def co(self, Y):  # :170-178
        x = self.xvec(Y)
        Axb = self.A @ x - self.b - self.y / self.sigma
        return float(self.c @ x + 0.5 * self.sigma * (Axb @ Axb))

    def cost(self, Y):  # :203-212
        x = self.xvec(Y)
        self.Axb = self.A @ x - self.b - self.y / self.sigma
        return float(self.c @ x + 0.5 * self.sigma * (self.Axb @ self.Axb))

    def accept(self, ok):
        pass

    def grad(self, Y):  # :214-226
        self.S = self.blocks(self.c + self.sigma * (self.At @ self.Axb))
        G, self.eG = [], []
        for i, Yi in enumerate(Y):
            Gi = 2 * (self.S[i] @ Yi)
            if i < self.nob:
                e = np.sum(Yi * Gi, axis=1, keepdims=True)
                Gi = Gi - Yi * e
            else:
                e = None
            self.eG.append(e)
            G.append(Gi)
        return Cells(G)

    def hess(self, Y, U):  # :228-247  (T = Y'*U of the p x n layout: T(a,b) = <Y_a, U_b>)
        YU = np.concatenate([_vec(Yi @ Ui.T) for Yi, Ui in zip(Y, U)])
        AyU = self.blocks(self.At @ (self.A @ YU))
        H = []
        for i, (Yi, Ui) in enumerate(zip(Y, U)):
            Hi = 2 * (self.S[i] @ Ui) + 4 * self.sigma * (AyU[i] @ Yi)
            if i < self.nob:
                Hi = Hi - Yi * np.sum(Yi * Hi, axis=1, keepdims=True) - Ui * self.eG[i]
            H.append(Hi)
        return Cells(H)


def _mb_normalize(Y, nob):
    return Cells([Yi / np.sqrt(np.sum(Yi * Yi, axis=1, keepdims=True)) if i < nob else Yi
                  for i, Yi in enumerate(Y)])


def _mb_line_search(co, Y, U, nob, literal_first_trial=False):
    """ManiSDP_multiblock.m:180-201.  The reference stacks [Y{i}; alpha*U{i}] (2p x n): because the
    rows of Y that U fills are zero, the Gram matrix -- hence the cost and every later iterate --
    equals that of Y + alpha*U, which is what is formed here so that the width stays p.
    `literal_first_trial`: line :184 builds the FIRST candidate from the empty nY (`[nY{i};
    alpha*U{i}]`), i.e. U alone; True reproduces that, False (default) starts from Y + U like the
    single-block drivers (ManiSDP_unitdiag.m:142)."""
    alpha = 1.0
    cost0 = co(Y)
    if literal_first_trial:
        nY = _mb_normalize(Cells([alpha * Ui for Ui in U]), nob)
    else:
        nY = _mb_normalize(Y + alpha * U, nob)
    k = 1
    while k <= 15 and co(nY) - cost0 > -1e-3:
        alpha *= 0.8
        nY = _mb_normalize(Y + alpha * U, nob)
        k += 1
    return nY


def ManiSDP_multiblock(At, b, c, K, options=None):
    """[X, obj, data] of src/primal/ManiSDP_multiblock.m:7 (K['s'] block orders, K['nob'] leading
    unit-diagonal blocks)."""
    o = dict(MB_DEFAULTS)
    o.update(options or {})
    o.setdefault("seed", 0)
    o.setdefault("verbose", False)
    n = [int(v) for v in np.atleast_1d(K["s"])]
    nb = len(n)
    nob = int(K.get("nob", 0))
    At = sp.csc_matrix(At)
    b = np.asarray(b.todense()).ravel() if sp.issparse(b) else np.asarray(b, dtype=np.float64).ravel()
    c = np.asarray(c.todense()).ravel() if sp.issparse(c) else np.asarray(c, dtype=np.float64).ravel()
    A = At.T.tocsr()
    p0 = o["p0"] if o["p0"] is not None else [1] * nb
    p = list(n)  # :33-39
    for i in range(nb):
        if n[i] >= o["min_facsize"]:
            p[i] = int(p0[i])
    sigma, gama = o["sigma0"], o["gama"]
    y = np.zeros(len(b))
    normb = 1 + np.linalg.norm(b)
    rng = np.random.default_rng(o["seed"])
    Y = o.get("Y0")
    if Y is not None:
        Y = Cells([np.array(Yi, dtype=np.float64) for Yi in Y])
        p = [Yi.shape[1] for Yi in Y]
    U = None
    data = dict(status=0, hv_count=0, tr_iters=0, fac_size=[])
    t0 = time.perf_counter()
    gap0 = pinf0 = dinf0 = None
    off = np.concatenate([[0], np.cumsum([v * v for v in n])]).astype(np.int64)
    for it in range(1, o["AL_maxiter"] + 1):
        data["fac_size"].append(list(p))
        prob = MultiblockProblem(At, b, c, n, p, nob, y, sigma)  # :61
        if Y is None:
            Y = prob.M.rand(rng)
        if U is not None:
            Y = _mb_line_search(prob.co, Y, U, nob, o.get("literal_first_trial", False))  # :62-64
        res = trustregions(prob, Y, maxiter=o["TR_maxiter"], maxinner=o["TR_maxinner"],
                           tolgradnorm=o["tolgradnorm"])  # :65
        Y = res.x
        data["hv_count"] += res.hv_count
        data["tr_iters"] += len(res.info) - 1
        gradnorm = res.info[-1].gradnorm
        X = [Yi @ Yi.T for Yi in Y]  # :67-72
        x = np.concatenate([_vec(Xi) for Xi in X])
        obj = float(c @ x)
        Axb = A @ x - b
        pinf = float(np.linalg.norm(Axb)) / normb
        y = y - sigma * Axb
        cy = c - At @ y
        by = float(b @ y)
        dinfs = np.zeros(nb)
        S, dS, vS = [], [], []
        for i in range(nb):  # :81-93
            Si = _mat(cy[off[i]:off[i + 1]], n[i])
            if i < nob:
                z = np.sum(X[i] * Si, axis=0)
                by += float(z.sum())
                Si = Si - np.diag(z)
            d, v = np.linalg.eigh(Si)
            S.append(Si)
            dS.append(d)
            vS.append(v)
            dinfs[i] = max(0.0, -d[0]) / (1 + abs(d[-1]))
        dinf = float(dinfs.max())
        gap = abs(obj - by) / (abs(by) + abs(obj) + 1)
        if o["verbose"]:
            print(f"Iter {it}, obj:{obj:0.8f}, gap:{gap:0.1e}, pinf:{pinf:0.1e}, dinf:{dinf:0.1e}, "
                  f"gradnorm:{gradnorm:0.1e}, p_max:{max(p)}, sigma:{sigma:0.3f}, "
                  f"time:{time.perf_counter()-t0:0.2f}s")
        eta = max(gap, pinf, dinf)
        if eta < o["tol"]:
            break
        if it % 50 == 0:  # :103-113
            if it > 100 and gap > gap0 and pinf > pinf0 and dinf > dinf0:
                data["status"] = 2
                break
            gap0, pinf0, dinf0 = gap, pinf, dinf
        Yl = list(Y.b)
        Ul = [None] * nb
        for i in range(nb):  # :114-153
            if n[i] < o["min_facsize"]:
                if o["line_search"] == 1:
                    Ul[i] = np.zeros_like(Yl[i])
                continue
            if p[i] > 1:
                Us, e, _ = np.linalg.svd(Yl[i], full_matrices=False)
                r = int(np.sum(e >= o["theta"] * e[0]))
                if r == 0:
                    r = 1
                if r < p[i]:
                    Yl[i] = Us[:, :r] * e[:r]
                    p[i] = r
            nneg = int(np.sum(dS[i] < 0))
            nne = max(min(nneg, o["delta"]), 1) if i < nob else min(nneg, o["delta"])
            if p[i] + nne > n[i]:
                nne = 0
            V = vS[i][:, :nne]
            if o["line_search"] == 1:
                Ul[i] = np.hstack([np.zeros((n[i], p[i])), V])
                Yl[i] = np.hstack([Yl[i], np.zeros((n[i], nne))])
            else:
                Yl[i] = np.hstack([Yl[i], o["alpha"] * V])
                if i < nob:
                    Yl[i] = Yl[i] / np.sqrt(np.sum(Yl[i] * Yl[i], axis=1, keepdims=True))
            p[i] += nne
        Y = Cells(Yl)
        U = Cells(Ul) if o["line_search"] == 1 else None
        if pinf < o["tau1"] * gradnorm:  # :154-158
            sigma = max(sigma / gama, o["sigma_min"])
        elif pinf > o["tau2"] * gradnorm:
            sigma = min(sigma * gama, o["sigma_max"])
    data.update(X=X, y=y, S=S, gap=gap, pinf=pinf, dinf=dinf, gradnorm=gradnorm,
                time=time.perf_counter() - t0, Y=Y, iters=it, obj=obj, sigma=sigma, dinfs=dinfs)
    if data["status"] == 0 and eta > o["tol"]:
        data["status"] = 1
    return X, obj, data


# --------------------------------------------------------------------------------------------
# dual approach: Riemannian ADMM on the SOS form (src/dual/ManiDSDP_unitdiag.m)
# --------------------------------------------------------------------------------------------
DUAL_DEFAULTS = dict(p0=None, ADMM_maxiter=300, gama=2, sigma0=1e-3, sigma_min=1e-3, sigma_max=1e7, tol=1e-8,
                     theta=1e-3, delta=8, alpha=0.1, tolgradnorm=1e-8, TR_maxinner=20, TR_maxiter=4, tau1=1e1,
                     tau2=1e2, line_search=0)  # :10-26


class DualProblem:
    """Closures of ManiDSDP_unitdiag.m:171-192 on the oblique manifold (rows of Y are unit vectors, S = Y Y').
    A: (m, n*n) PSD part of the constraint matrix, B: (m, K.f) free part, iA = (diag(dAAt) \\ A)'."""

    def __init__(self, A, B, b, c, cf, n, p, dAAt, x, w, sigma):
        self.A, self.B, self.b, self.c, self.cf, self.n = A, B, b, c, cf, n
        self.iAt = sp.diags(1.0 / dAAt) @ A  # iA' = D^{-1} A   (:43)
        self.bA = self.iAt.T @ b  # :44
        self.x, self.w, self.sigma = x, w, sigma
        self.M = ObliqueT(n, p)

    def co(self, Y):  # :149-156
        sc = _vec(Y @ Y.T) - self.c
        y = self.iAt @ sc
        As = self.A.T @ y - sc - self.x / self.sigma
        Af = self.B.T @ y - self.cf - self.w / self.sigma
        return float(self.b @ y + 0.5 * self.sigma * (As @ As + Af @ Af))

    def cost(self, Y):  # :171-178
        sc = _vec(Y @ Y.T) - self.c
        y = self.iAt @ sc
        self.As = self.A.T @ y - sc - self.x / self.sigma
        Af = self.B.T @ y - self.cf - self.w / self.sigma
        return float(self.b @ y + 0.5 * self.sigma * (self.As @ self.As + Af @ Af))

    def accept(self, ok):
        pass

    def grad(self, Y):  # :180-184
        self.X = _mat(self.bA - self.sigma * self.As, self.n)
        eG = 2 * (self.X @ Y)
        self.YeG = np.sum(Y * eG, axis=1, keepdims=True)
        return eG - Y * self.YeG

    def hess(self, Y, U):  # :186-191  (row layout: every product transposed)
        s = self.sigma
        YU = Y @ U.T  # (Y'*U)(a,b) = <Y_a, U_b>
        yAU = _mat(self.A.T @ (self.iAt @ _vec(YU)), self.n)
        eH = 2 * (self.X @ U) - 4 * s * (yAU.T @ Y) + 2 * s * (Y @ (U.T @ Y) + U @ (Y.T @ Y))
        return eH - Y * np.sum(Y * eH, axis=1, keepdims=True) - U * self.YeG


def ManiDSDP_unitdiag(A, b, c, K, options=None):
    """[X, obj, data] of src/dual/ManiDSDP_unitdiag.m:8.  A: (m, K.f + n*n), c: (K.f + n*n,)."""
    o = dict(DUAL_DEFAULTS)
    o.update(options or {})
    o.setdefault("seed", 0)
    o.setdefault("verbose", False)
    n = int(K["s"])
    nf = int(K.get("f", 0))
    A = sp.csr_matrix(A)
    b = np.asarray(b, dtype=np.float64).ravel()
    c = np.asarray(c.todense()).ravel() if sp.issparse(c) else np.asarray(c, dtype=np.float64).ravel()
    if o["p0"] is None:
        o["p0"] = int(math.ceil(math.log(len(b))))  # :11
    normc = 1 + np.linalg.norm(c)  # :33
    B = A[:, :nf].tocsr()
    A = A[:, nf:].tocsr()
    cf, c = c[:nf], c[nf:]
    dAAt = np.asarray(o["dAAt"], dtype=np.float64).ravel() if o.get("dAAt") is not None else \
        np.asarray(A.multiply(A).sum(axis=1)).ravel()  # :40
    p = int(o["p0"])
    sigma, gama = o["sigma0"], o["gama"]
    x = np.zeros(n * n)
    w = np.zeros(nf)
    rng = np.random.default_rng(o["seed"])
    Y = o.get("Y0")
    if Y is not None:
        Y = np.array(Y, dtype=np.float64)
        p = Y.shape[1]
    U = None
    data = dict(status=0, hv_count=0, tr_iters=0, fac_size=[], seta=[])
    t0 = time.perf_counter()
    gap0 = pinf0 = dinf0 = None
    for it in range(1, o["ADMM_maxiter"] + 1):
        data["fac_size"].append(p)
        prob = DualProblem(A, B, b, c, cf, n, p, dAAt, x, w, sigma)
        if Y is None:
            Y = prob.M.rand(rng)
        if U is not None:
            Y = _line_search("unitdiag", prob.co, Y, U)  # :66-68, :158-169
        res = trustregions(prob, Y, maxiter=o["TR_maxiter"], maxinner=o["TR_maxinner"],
                           tolgradnorm=o["tolgradnorm"])
        Y = res.x
        data["hv_count"] += res.hv_count
        data["tr_iters"] += len(res.info) - 1
        gradnorm = res.info[-1].gradnorm
        S = Y @ Y.T  # :71-76
        sc = _vec(S) - c
        y = prob.iAt @ sc
        As = A.T @ y - sc
        Af = B.T @ y - cf
        pinf = (np.linalg.norm(As) + np.linalg.norm(Af)) / normc
        by = float(b @ y)
        x = x - sigma * As  # :78-79
        w = w - sigma * Af
        eX = _mat(x + prob.bA, n)
        z = np.sum(S * eX, axis=0)
        X = eX - np.diag(z)
        dX, vX = np.linalg.eigh(X)
        obj = float(c @ _vec(eX) + cf @ w + z.sum())  # :86
        dinf = max(0.0, -dX[0]) / (1 + abs(dX[-1]))
        gap = abs(obj - by) / (1 + abs(obj) + abs(by))
        Us, e, _ = np.linalg.svd(Y, full_matrices=False)
        r = int(np.sum(e > o["theta"] * e[0]))  # :91 (strict)
        if o["verbose"]:
            print(f"Iter {it}, obj:{obj:0.8f}, gap:{gap:0.1e}, pinf:{pinf:0.1e}, dinf:{dinf:0.1e}, "
                  f"gradnorm:{gradnorm:0.1e}, r:{r}, p:{p}, sigma:{sigma:0.3f}, "
                  f"time:{time.perf_counter()-t0:0.2f}s")
        eta = max(gap, pinf, dinf)
        data["seta"].append(eta)
        if eta < o["tol"]:
            break
        if it % 50 == 0:
            if it > 100 and gap > gap0 and pinf > pinf0 and dinf > dinf0:
                data["status"] = 2
                break
            gap0, pinf0, dinf0 = gap, pinf, dinf
        if r <= p - 1:  # :112-115
            Y, p = Us[:, :r] * e[:r], r
        nne = max(min(int(np.sum(dX < 0)), o["delta"]), 1)  # :116
        Y, U = _escape("unitdiag", o, Y, vX, nne)
        p += nne
        if pinf < o["tau1"] * gradnorm:
            sigma = max(sigma / gama, o["sigma_min"])
        elif pinf > o["tau2"] * gradnorm:
            sigma = min(sigma * gama, o["sigma_max"])
    data.update(X=X, y=y, S=S, w=w, x=x, gap=gap, pinf=pinf, dinf=dinf, gradnorm=gradnorm,
                time=time.perf_counter() - t0, Y=Y, iters=it, obj=obj, sigma=sigma)
    if data["status"] == 0 and eta > o["tol"]:
        data["status"] = 1
    return X, obj, data
