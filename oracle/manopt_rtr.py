"""ORACLE (test infrastructure, not product code) -- Riemannian trust regions + truncated CG.

CPU restatement (NumPy FP64) of the two Manopt 7.0 routines on ManiSDP's hot path:

  * `tCG`            follows manopt7.0/manopt/solvers/trustregions/tCG.m:95-292
  * `trustregions`   follows manopt7.0/manopt/solvers/trustregions/trustregions.m:340-372 (defaults),
                     :395-417 (initialisation), :441-767 (main loop), with useRand = false, no
                     preconditioner (getPrecon.m:62-66 is the identity), no hooks / statsfun.

and of the three manifold structs the solvers use:

  * `ObliqueT`   the inline transposed oblique factory, src/primal/ManiSDP_unitdiag.m:173-198
  * `Sphere`     manopt7.0/manopt/manifolds/sphere/spherefactory.m:83-115,220-232
  * `Euclid`     manopt7.0/manopt/manifolds/euclidean/euclideanfactory.m:49-94

LAYOUT: every point / tangent vector in this oracle is an (n, p) array whose ROW i belongs to
vertex i (row i of the SDP factor).  The unit-diagonal reference drivers store the transpose
(p x n, unit COLUMNS); formulas are transposed accordingly, arithmetic is unchanged.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

EPS = float(np.finfo(np.float64).eps)


# --------------------------------------------------------------------------------------------
# manifolds
# --------------------------------------------------------------------------------------------
class ObliqueT:
    """Rows of Y on the unit sphere S^{p-1} (ManiSDP_unitdiag.m:173-198, transposed layout)."""

    name = "oblique"

    def __init__(self, n, p):
        self.n, self.p = n, p

    def dim(self):
        return (self.p - 1) * self.n

    def typicaldist(self):
        return math.pi * math.sqrt(self.n)  # :179

    def inner(self, x, a, b):
        return float(np.vdot(a, b))  # :176

    def norm(self, x, a):
        return float(np.linalg.norm(a))  # :177

    def proj(self, x, u):
        return u - x * np.sum(x * u, axis=1, keepdims=True)  # :180-181

    tangent = proj

    def retr(self, x, d):
        y = x + d
        return y / np.sqrt(np.sum(y * y, axis=1, keepdims=True))  # :184-187

    def rand(self, rng):
        x = rng.standard_normal((self.n, self.p))
        return x / np.sqrt(np.sum(x * x, axis=1, keepdims=True))  # :194-197

    def zerovec(self, x):
        return np.zeros_like(x)


class Sphere:
    """Unit Frobenius-norm n x p matrices (spherefactory.m)."""

    name = "sphere"

    def __init__(self, n, p):
        self.n, self.p = n, p

    def dim(self):
        return self.n * self.p - 1  # :83

    def typicaldist(self):
        return math.pi  # :111

    def inner(self, x, a, b):
        return float(np.vdot(a, b))  # :85

    def norm(self, x, a):
        return float(np.linalg.norm(a))  # :87

    def proj(self, x, d):
        return d - x * float(np.vdot(x, d))  # :113

    tangent = proj

    def retr(self, x, d):
        y = x + d
        return y / np.linalg.norm(y)  # :220-232

    def rand(self, rng):
        x = rng.standard_normal((self.n, self.p))
        return x / np.linalg.norm(x)  # :249-254

    def zerovec(self, x):
        return np.zeros_like(x)


class Euclid:
    """Flat R^{n x p} (euclideanfactory.m:49-94)."""

    name = "euclid"

    def __init__(self, n, p):
        self.n, self.p = n, p

    def dim(self):
        return self.n * self.p

    def typicaldist(self):
        return math.sqrt(self.n * self.p)  # :59

    def inner(self, x, a, b):
        return float(np.vdot(a, b))

    def norm(self, x, a):
        return float(np.linalg.norm(a))

    def proj(self, x, d):
        return d

    tangent = proj

    def retr(self, x, d):
        return x + d

    def rand(self, rng):
        return rng.standard_normal((self.n, self.p))

    def zerovec(self, x):
        return np.zeros_like(x)


# --------------------------------------------------------------------------------------------
# truncated CG
# --------------------------------------------------------------------------------------------
def tCG(M, hess, x, grad, Delta, maxinner, mininner=1, kappa=0.1, theta=1.0):
    """Steihaug-Toint truncated CG from eta0 = 0 with the identity preconditioner.

    hess(u) -> Hess f(x)[u].  Returns (eta, Heta, inner_it, stop_tCG) with stop codes
    1 negative curvature, 2 exceeded trust region, 3 kappa (linear), 4 theta (superlinear),
    5 maximum inner iterations, 6 model increased  (tCG.m:167-171 of the header / :183-257)."""
    inner = lambda a, b: M.inner(x, a, b)
    eta = M.zerovec(x)
    Heta = M.zerovec(x)
    r = grad  # tCG.m:106
    e_Pe = 0.0
    r_r = inner(r, r)
    norm_r = math.sqrt(r_r)
    norm_r0 = norm_r
    z = r  # identity preconditioner, tCG.m:120
    z_r = inner(z, r)
    d_Pd = z_r
    mdelta = z  # tCG.m:131
    e_Pd = 0.0
    model_value = 0.0  # tCG.m:151-152
    stop = 5
    j = 0
    for j in range(1, maxinner + 1):
        Hmdelta = hess(mdelta)  # tCG.m:163
        d_Hd = inner(mdelta, Hmdelta)  # :166
        alpha = z_r / d_Hd if d_Hd != 0.0 else math.copysign(math.inf, z_r)  # :170
        e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd  # :173
        if d_Hd <= 0 or e_Pe_new >= Delta**2:  # :183
            tau = (-e_Pd + math.sqrt(e_Pd * e_Pd + d_Pd * (Delta**2 - e_Pe))) / d_Pd  # :188
            eta = eta - tau * mdelta  # :192
            Heta = Heta - tau * Hmdelta  # :196
            stop = 1 if d_Hd <= 0 else 2  # :205-209
            break
        e_Pe = e_Pe_new  # :214
        new_eta = eta - alpha * mdelta  # :215
        new_Heta = Heta - alpha * Hmdelta  # :220
        new_model_value = inner(new_eta, grad) + 0.5 * inner(new_eta, new_Heta)  # :227 (model_fun :150)
        if new_model_value >= model_value:  # :228
            stop = 6
            break
        eta, Heta, model_value = new_eta, new_Heta, new_model_value  # :233-235
        r = r - alpha * Hmdelta  # :238
        r_r = inner(r, r)  # :241
        norm_r = math.sqrt(r_r)
        if j >= mininner and norm_r <= norm_r0 * min(norm_r0**theta, kappa):  # :249
            stop = 3 if kappa < norm_r0**theta else 4  # :251-255
            break
        z = r  # :260-264
        zold_rold = z_r  # :267
        z_r = inner(z, r)  # :269
        beta = z_r / zold_rold  # :272
        mdelta = M.tangent(x, z + beta * mdelta)  # :273, :283
        e_Pd = beta * (e_Pd + alpha * d_Pd)  # :286
        d_Pd = z_r + beta * beta * d_Pd  # :287
    return eta, Heta, j, stop


# --------------------------------------------------------------------------------------------
# trust regions
# --------------------------------------------------------------------------------------------
@dataclass
class TRInfo:
    iter: int
    cost: float
    gradnorm: float
    Delta: float
    rho: float = math.inf
    accepted: bool = True
    numinner: int = 0
    stop_inner: int = 0
    stepsize: float = math.nan


@dataclass
class TRResult:
    x: np.ndarray
    cost: float
    info: list = field(default_factory=list)
    hv_count: int = 0


def trustregions(problem, x, maxiter=1000, maxinner=None, tolgradnorm=1e-6, mininner=1,
                 kappa=0.1, theta=1.0, rho_prime=0.1, rho_regularization=1e3,
                 Delta_bar=None, Delta0=None):
    """problem: object with .M (manifold), .cost(x)->f (must be called before .grad at the same
    point: the reference's closures share state that way), .grad(x)->G,
    .hess(x, u)->H, and .accept(ok) to tell a stateful problem whether the last cost() point was
    kept (needed only because this oracle keeps per-point caches explicit; see problem classes).

    Returns TRResult.  Restates trustregions.m:395-767."""
    M = problem.M
    if maxinner is None:
        maxinner = M.dim()
    if Delta_bar is None:
        Delta_bar = M.typicaldist()  # :361-366
    if Delta0 is None:
        Delta0 = Delta_bar / 8  # :368-370
    res = TRResult(x=x, cost=math.nan)
    k = 0
    fx = problem.cost(x)  # getCostGrad, :405
    problem.accept(True)
    fgradx = problem.grad(x)
    norm_grad = M.norm(x, fgradx)
    Delta = Delta0
    res.info.append(TRInfo(0, fx, norm_grad, Delta))
    while True:
        # stoppingcriterion.m:50-72 (tolcost/maxtime never set by ManiSDP)
        if norm_grad < tolgradnorm:
            break
        if k >= maxiter:
            break
        nhv = [0]

        def hess(u, _x=x):
            nhv[0] += 1
            return problem.hess(_x, u)

        eta, Heta, numit, stop_inner = tCG(M, hess, x, fgradx, Delta, maxinner, mininner, kappa, theta)
        res.hv_count += nhv[0]
        norm_eta = M.norm(x, eta)  # :531
        x_prop = M.retr(x, eta)  # :540
        fx_prop = problem.cost(x_prop)  # :544
        rhonum = fx - fx_prop  # :548
        rhoden = -M.inner(x, eta, fgradx + 0.5 * Heta)  # :549-550
        rho_reg = max(1.0, abs(fx)) * EPS * rho_regularization  # :579
        rhonum += rho_reg
        rhoden += rho_reg
        model_decreased = rhoden >= 0  # :613
        rho = rhonum / rhoden if rhoden != 0 else math.nan  # :621
        if rho < 0.25 or not model_decreased or math.isnan(rho):  # :653
            Delta = Delta / 4
        elif rho > 0.75 and stop_inner in (1, 2):  # :667
            Delta = min(2 * Delta, Delta_bar)
        if model_decreased and rho > rho_prime:  # :688
            accept = True
            problem.accept(True)
            x = x_prop
            fx = fx_prop
            fgradx = problem.grad(x)  # :718
            norm_grad = M.norm(x, fgradx)
        else:
            accept = False
            problem.accept(False)
        k += 1
        res.info.append(TRInfo(k, fx, norm_grad, Delta, rho, accept, numit, stop_inner, norm_eta))
    res.x = x
    res.cost = fx
    return res


# --------------------------------------------------------------------------------------------
# product manifold of the multi-block driver
# --------------------------------------------------------------------------------------------
class Cells:
    """A MATLAB cell array of matrices with the arithmetic `lincombc.cpp:21-66` provides
    (a1*u1, a1*u1 + a2*u2), so that `tCG` / `trustregions` above run unchanged on a product point.
    Block i is an (n_i, p_i) array (row layout, see the module docstring)."""

    __array_ufunc__ = None  # NumPy scalars on the left defer to __rmul__

    def __init__(self, blocks):
        self.b = list(blocks)

    def __len__(self):
        return len(self.b)

    def __getitem__(self, i):
        return self.b[i]

    def __iter__(self):
        return iter(self.b)

    def __add__(self, o):
        return Cells([x + y for x, y in zip(self.b, o.b)])

    def __sub__(self, o):
        return Cells([x - y for x, y in zip(self.b, o.b)])

    def __mul__(self, a):
        return Cells([x * float(a) for x in self.b])

    __rmul__ = __mul__

    def __neg__(self):
        return Cells([-x for x in self.b])

    def copy(self):
        return Cells([x.copy() for x in self.b])


class MultiBlock:
    """src/basicfunction/multiblockmanifold.m:1-42: the first `nob` blocks are transposed oblique
    manifolds (unit rows here), the others Euclidean.

    The manifold operations are MEX files in the reference (src/C-files/{innerc,projc,retrc,
    lincombc,randc,zerovecc}.cpp).  Two of the committed sources are older than the .m call sites
    (retrc.cpp takes 3 arguments and normalises every block, multiblockmanifold.m:24 passes 4;
    projc.cpp:34-43 subtracts ONE scalar <X_i,U_i> per block instead of one per unit vector).  This
    restatement follows the .m call sites and the closures of ManiSDP_multiblock.m:214-247, which
    use the per-vector form (`Y{i}.*sum(Y{i}.*H{i})`); on the tangent inputs tCG passes to M.proj
    (tCG.m:273,283) the two forms of projc agree to rounding because every per-vector product is
    already zero."""

    name = "multiblock"

    def __init__(self, pset, nset, nob):
        self.p, self.n, self.nob = [int(v) for v in pset], [int(v) for v in nset], int(nob)

    def dim(self):  # :3
        return sum((p - 1) * n for p, n in zip(self.p[:self.nob], self.n[:self.nob])) + \
            sum(p * n for p, n in zip(self.p[self.nob:], self.n[self.nob:]))

    def typicaldist(self):  # :11-15
        return math.sqrt(math.pi * sum(self.n[:self.nob]) +
                         sum(p * n for p, n in zip(self.p[self.nob:], self.n[self.nob:])))

    def inner(self, x, a, b):  # innerc.cpp:20-32
        return float(sum(np.vdot(u, v) for u, v in zip(a, b)))

    def norm(self, x, a):
        return math.sqrt(self.inner(x, a, a))

    def proj(self, x, u):  # projc.cpp:19-56 (per unit vector, see the class docstring)
        return Cells([ui - xi * np.sum(xi * ui, axis=1, keepdims=True) if i < self.nob else ui.copy()
                      for i, (xi, ui) in enumerate(zip(x, u))])

    tangent = proj

    def retr(self, x, d):  # retrc.cpp:24-47 with the nob argument of multiblockmanifold.m:24
        out = []
        for i, (xi, di) in enumerate(zip(x, d)):
            y = xi + di
            out.append(y / np.sqrt(np.sum(y * y, axis=1, keepdims=True)) if i < self.nob else y)
        return Cells(out)

    def rand(self, rng):  # randc.cpp:52-80
        out = []
        for i, (p, n) in enumerate(zip(self.p, self.n)):
            x = rng.standard_normal((n, p))
            out.append(x / np.sqrt(np.sum(x * x, axis=1, keepdims=True)) if i < self.nob else x)
        return Cells(out)

    def zerovec(self, x):
        return Cells([np.zeros_like(xi) for xi in x])
