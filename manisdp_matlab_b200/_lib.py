"""ctypes binding of libmanisdp_b200.so (include/manisdp_b200.h).

This is the Python stand-in for the MATLAB MEX gateway (matlab/manisdp_mex.cpp): it passes plain pointers and sizes
to the same C ABI.  There is NO CPU fallback here: if the shared library is missing, or no B200 is visible, every
compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmanisdp_b200.so")

ONLYUNITDIAG, UNITDIAG, UNITTRACE, GENERAL, MULTIBLOCK, DUAL_UNITDIAG = 0, 1, 2, 3, 4, 5
LAYOUT_ROWS, LAYOUT_COLS = 0, 1
SLOT_Y, SLOT_YPROP, SLOT_G, SLOT_ETA, SLOT_R, SLOT_D, SLOT_HD, SLOT_U, SLOT_H = range(9)
KIND_NAMES = {"onlyunitdiag": ONLYUNITDIAG, "unitdiag": UNITDIAG, "unittrace": UNITTRACE, "general": GENERAL,
              "multiblock": MULTIBLOCK, "dual_unitdiag": DUAL_UNITDIAG}

_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)


class Problem(C.Structure):
    _fields_ = [("kind", C.c_int32), ("device", C.c_int32), ("n", C.c_int64), ("m", C.c_int64),
                ("C_jc", _u64p), ("C_ir", _u64p), ("C_pr", _f64p),
                ("At_jc", _u64p), ("At_ir", _u64p), ("At_pr", _f64p),
                ("b", _f64p), ("c_ir", _u64p), ("c_pr", _f64p), ("c_nnz", C.c_int64),
                ("rank", C.c_int32), ("world", C.c_int32), ("row_begin", C.c_int64), ("row_end", C.c_int64),
                ("nccl_unique_id", C.c_void_p), ("force_mode", C.c_int32), ("shard_layout", C.c_int32),
                ("nblocks", C.c_int32), ("nob", C.c_int32), ("block_sizes", C.POINTER(C.c_int64)),
                ("dAAt", _f64p), ("nfree", C.c_int64), ("B_jc", _u64p), ("B_ir", _u64p), ("B_pr", _f64p),
                ("cf", _f64p)]


class TrOptions(C.Structure):
    _fields_ = [("maxiter", C.c_int32), ("maxinner", C.c_int32), ("mininner", C.c_int32), ("use_graph", C.c_int32),
                ("tolgradnorm", C.c_double), ("kappa", C.c_double), ("theta", C.c_double),
                ("rho_prime", C.c_double), ("rho_regularization", C.c_double), ("Delta_bar", C.c_double),
                ("Delta0", C.c_double)]


class TrInfo(C.Structure):
    _fields_ = [("cost", C.c_double), ("gradnorm", C.c_double), ("Delta", C.c_double), ("seconds", C.c_double),
                ("hv_count", C.c_int64), ("iters", C.c_int32), ("accepted", C.c_int32), ("stop_reason", C.c_int32),
                ("reserved", C.c_int32)]


class TrIter(C.Structure):
    _fields_ = [("cost", C.c_double), ("gradnorm", C.c_double), ("Delta", C.c_double), ("rho", C.c_double),
                ("stepsize", C.c_double), ("iter", C.c_int32), ("numinner", C.c_int32), ("stop_inner", C.c_int32),
                ("accepted", C.c_int32)]


class KktInfo(C.Structure):
    _fields_ = [("obj", C.c_double), ("by", C.c_double), ("pinf", C.c_double), ("dinf", C.c_double),
                ("gap", C.c_double), ("lam_min", C.c_double), ("lam_max", C.c_double), ("z_sum", C.c_double),
                ("nneg", C.c_int32), ("eig_iters", C.c_int32), ("eig_resid", C.c_double),
                ("eig_converged", C.c_int32), ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_local", C.c_int64), ("m", C.c_int64), ("p", C.c_int64), ("ld", C.c_int64),
                ("nnzC", C.c_int64), ("nnzA", C.c_int64), ("kind", C.c_int32), ("s_mode", C.c_int32),
                ("a_mode", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("hv_total", C.c_int64),
                ("launches_total", C.c_int64), ("bytes_per_hv", C.c_double), ("flops_per_hv", C.c_double)]


# every symbol include/manisdp_b200.h declares: (name, restype, argtypes)
_H = C.c_void_p
SIGNATURES = {
    "manisdp_create": (C.c_int, [C.POINTER(_H), C.POINTER(Problem)]),
    "manisdp_destroy": (C.c_int, [_H]),
    "manisdp_last_error": (C.c_char_p, [_H]),
    "manisdp_version": (C.c_int, []),
    "manisdp_set_Y": (C.c_int, [_H, _f64p, C.c_int64, C.c_int32]),
    "manisdp_get_Y": (C.c_int, [_H, _f64p, C.c_int32]),
    "manisdp_get_p": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "manisdp_rand_Y": (C.c_int, [_H, C.c_int64, C.c_uint64]),
    "manisdp_set_dual": (C.c_int, [_H, _f64p, C.c_double]),
    "manisdp_get_dual": (C.c_int, [_H, _f64p, C.POINTER(C.c_double)]),
    "manisdp_slot_set": (C.c_int, [_H, C.c_int32, _f64p, C.c_int32]),
    "manisdp_slot_get": (C.c_int, [_H, C.c_int32, _f64p, C.c_int32]),
    "manisdp_cost": (C.c_int, [_H, C.POINTER(C.c_double)]),
    "manisdp_grad": (C.c_int, [_H, C.POINTER(C.c_double)]),
    "manisdp_hess": (C.c_int, [_H]),
    "manisdp_hess_bench": (C.c_int, [_H, C.c_int32, C.POINTER(C.c_double)]),
    "manisdp_vec_bench": (C.c_int, [_H, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "manisdp_retract": (C.c_int, [_H, C.c_int32, C.c_int32]),
    "manisdp_project": (C.c_int, [_H, C.c_int32, C.c_int32]),
    "manisdp_tr_solve": (C.c_int, [_H, C.POINTER(TrOptions), C.POINTER(TrInfo)]),
    "manisdp_tr_log": (C.c_int, [_H, C.POINTER(TrIter), C.c_int32, C.POINTER(C.c_int32)]),
    "manisdp_kkt": (C.c_int, [_H, C.c_int32, C.c_double, C.c_int32, C.POINTER(KktInfo)]),
    "manisdp_get_eigs": (C.c_int, [_H, _f64p, _f64p, C.c_int32]),
    "manisdp_rank_cut": (C.c_int, [_H, C.c_double, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "manisdp_escape": (C.c_int, [_H, C.c_int32, C.c_double, C.c_int32]),
    "manisdp_line_search": (C.c_int, [_H, C.POINTER(C.c_double)]),
    "manisdp_set_sigma": (C.c_int, [_H, C.c_double]),
    "manisdp_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "manisdp_get_stats": (C.c_int, [_H, C.POINTER(Stats)]),
    "manisdp_test_sym_eig": (C.c_int, [_f64p, C.c_int32, _f64p, _f64p]),
    "manisdp_col_split": (C.c_int, [_H]),
    "manisdp_col_merge": (C.c_int, [_H]),
    "manisdp_group_create": (C.c_int, [C.POINTER(_H), C.POINTER(Problem), C.c_int32, C.POINTER(C.c_int32)]),
    "manisdp_group_destroy": (C.c_int, [_H]),
    "manisdp_group_size": (C.c_int, [_H]),
    "manisdp_group_last_error": (C.c_char_p, [_H]),
    "manisdp_group_set_Y": (C.c_int, [_H, _f64p, C.c_int64, C.c_int32]),
    "manisdp_group_rand_Y": (C.c_int, [_H, C.c_int64, C.c_uint64]),
    "manisdp_group_get_Y": (C.c_int, [_H, _f64p, C.c_int32]),
    "manisdp_group_get_stats": (C.c_int, [_H, C.POINTER(Stats)]),
    "manisdp_group_cost": (C.c_int, [_H, _f64p]),
    "manisdp_group_tr_solve": (C.c_int, [_H, C.POINTER(TrOptions), C.POINTER(TrInfo)]),
    "manisdp_group_kkt": (C.c_int, [_H, C.c_int32, C.c_double, C.c_int32, C.POINTER(KktInfo)]),
    "manisdp_group_rank_cut": (C.c_int, [_H, C.c_double, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "manisdp_group_escape": (C.c_int, [_H, C.c_int32, C.c_double, C.c_int32]),
    "manisdp_group_line_search": (C.c_int, [_H, _f64p]),
    "manisdp_get_index_split": (C.c_int, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int64,
                                          C.POINTER(C.c_int64)]),
    "manisdp_dual_get_state": (C.c_int, [_H, _f64p, _f64p]),
    "manisdp_dual_set_state": (C.c_int, [_H, _f64p, _f64p]),
    "manisdp_mb_set_Y": (C.c_int, [_H, _f64p, C.POINTER(C.c_int64)]),
    "manisdp_mb_get_Y": (C.c_int, [_H, _f64p]),
    "manisdp_mb_get_widths": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "manisdp_mb_rand_Y": (C.c_int, [_H, C.POINTER(C.c_int64), C.c_uint64]),
    "manisdp_mb_kkt": (C.c_int, [_H, C.c_int32, C.POINTER(KktInfo), _f64p, C.POINTER(C.c_int32)]),
    "manisdp_mb_get_block_eigs": (C.c_int, [_H, C.c_int32, _f64p, _f64p]),
    "manisdp_mb_update": (C.c_int, [_H, C.c_double, C.c_int32, C.c_double, C.c_int32, C.c_int32,
                                    C.POINTER(C.c_int64)]),
}

_lib = None


def load():
    """dlopen the engine (once).  Raises if the library was not built: the product has no other path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m manisdp_matlab_b200.build` (or __graft_entry__.build()). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class EngineError(RuntimeError):
    pass


def _as_u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _pu(a):
    return a.ctypes.data_as(_u64p)


def _pf(a):
    return a.ctypes.data_as(_f64p)


class Handle:
    """Thin object wrapper over manisdp_t*; every method maps 1:1 onto a C-ABI call."""

    def __init__(self, kind, n, *, C_csc=None, At=None, b=None, c=None, device=0, rank=0, world=1,
                 row_begin=0, row_end=None, nccl_id=None, force_mode=0, layout="rows", block_sizes=None, nob=0,
                 dAAt=None, B=None, cf=None):
        import scipy.sparse as sp

        self.lib = load()
        self._h = _H()
        pb = Problem()
        pb.kind = KIND_NAMES[kind] if isinstance(kind, str) else int(kind)
        pb.device = device
        pb.n = n
        pb.rank, pb.world = rank, world
        pb.row_begin = row_begin
        pb.row_end = n if row_end is None else row_end
        pb.force_mode = force_mode
        pb.shard_layout = {"rows": 0, "cols": 1}[layout]
        self.layout, self.rank, self.world = layout, rank, world
        keep = []
        if pb.kind == MULTIBLOCK:  # K.s, K.nob of ManiSDP_multiblock.m:7
            bs = np.ascontiguousarray(block_sizes, dtype=np.int64)
            assert int(bs.sum()) == n, "n must equal sum(block_sizes)"
            keep.append(bs)
            pb.nblocks, pb.nob = len(bs), int(nob)
            pb.block_sizes = bs.ctypes.data_as(C.POINTER(C.c_int64))
            self.block_sizes, self.nob = [int(v) for v in bs], int(nob)
        if pb.kind == ONLYUNITDIAG:
            Cm = sp.csc_matrix(C_csc)
            Cm.sort_indices()
            jc, ir, pr = _as_u64(Cm.indptr), _as_u64(Cm.indices), _as_f64(Cm.data)
            keep += [jc, ir, pr]
            pb.C_jc, pb.C_ir, pb.C_pr = _pu(jc), _pu(ir), _pf(pr)
            pb.m = 0
        else:
            Atm = sp.csc_matrix(At)
            Atm.sort_indices()
            jc, ir, pr = _as_u64(Atm.indptr), _as_u64(Atm.indices), _as_f64(Atm.data)
            bb = _as_f64(np.asarray(b.todense()).ravel() if sp.issparse(b) else np.asarray(b).ravel())
            keep += [jc, ir, pr, bb]
            pb.At_jc, pb.At_ir, pb.At_pr, pb.b = _pu(jc), _pu(ir), _pf(pr), _pf(bb)
            pb.m = Atm.shape[1]
            if sp.issparse(c):
                cc = sp.csc_matrix(c.reshape(-1, 1))
                cir, cpr = _as_u64(cc.indices), _as_f64(cc.data)
                keep += [cir, cpr]
                pb.c_ir, pb.c_pr, pb.c_nnz = _pu(cir), _pf(cpr), len(cpr)
            else:
                cpr = _as_f64(np.asarray(c).ravel())
                keep += [cpr]
                pb.c_ir, pb.c_pr, pb.c_nnz = None, _pf(cpr), len(cpr)
        self.nfree = 0
        if pb.kind == DUAL_UNITDIAG:  # ManiDSDP_unitdiag.m:35-44: free part B (m x K.f), its cost cf, options.dAAt
            if dAAt is not None:
                dd = _as_f64(np.asarray(dAAt).ravel())
                assert dd.shape == (pb.m,)
                keep.append(dd)
                pb.dAAt = _pf(dd)
            if B is not None and B.shape[1] > 0:
                Bm = sp.csc_matrix(B)
                Bm.sort_indices()
                bjc, bir, bpr = _as_u64(Bm.indptr), _as_u64(Bm.indices), _as_f64(Bm.data)
                cff = _as_f64(np.asarray(cf).ravel())
                assert Bm.shape == (pb.m, len(cff))
                keep += [bjc, bir, bpr, cff]
                pb.nfree, pb.B_jc, pb.B_ir, pb.B_pr, pb.cf = Bm.shape[1], _pu(bjc), _pu(bir), _pf(bpr), _pf(cff)
                self.nfree = Bm.shape[1]
        if nccl_id is not None:
            idbuf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
            keep.append(idbuf)
            pb.nccl_unique_id = C.cast(idbuf, C.c_void_p)
        rc = self.lib.manisdp_create(C.byref(self._h), C.byref(pb))
        if rc != 0:
            msg = self.lib.manisdp_last_error(None)
            raise EngineError(f"manisdp_create failed ({rc}): {msg.decode() if msg else ''}")
        self.n = n
        self.n_local = pb.row_end - pb.row_begin
        self.m = pb.m
        self.kind = pb.kind

    # -- plumbing
    def _ck(self, rc, what):
        if rc != 0:
            msg = self.lib.manisdp_last_error(self._h)
            raise EngineError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if self._h:
            self.lib.manisdp_destroy(self._h)
            self._h = _H()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- state
    @property
    def p(self):
        v = C.c_int64()
        self._ck(self.lib.manisdp_get_p(self._h, C.byref(v)), "get_p")
        return v.value

    def set_Y(self, Y):
        """Y: (n_local, p) array, one row per vertex."""
        Y = _as_f64(Y)
        assert Y.ndim == 2 and Y.shape[0] == self.n_local
        self._ck(self.lib.manisdp_set_Y(self._h, _pf(Y), Y.shape[1], LAYOUT_ROWS), "set_Y")

    def set_Y_cols(self, Yf):
        """Yf: Fortran-ordered (n, p) array (MATLAB's layout in ManiSDP.m / ManiSDP_unittrace.m)."""
        Yf = np.asfortranarray(Yf, dtype=np.float64)
        self._ck(self.lib.manisdp_set_Y(self._h, Yf.ctypes.data_as(_f64p), Yf.shape[1], LAYOUT_COLS), "set_Y")

    def get_Y(self):
        out = np.empty((self.n_local, self.p))
        self._ck(self.lib.manisdp_get_Y(self._h, _pf(out), LAYOUT_ROWS), "get_Y")
        return out

    def get_Y_cols(self):
        out = np.empty((self.n_local, self.p), order="F")
        self._ck(self.lib.manisdp_get_Y(self._h, out.ctypes.data_as(_f64p), LAYOUT_COLS), "get_Y")
        return out

    def rand_Y(self, p, seed=0):
        self._ck(self.lib.manisdp_rand_Y(self._h, p, seed), "rand_Y")

    def set_dual(self, y, sigma):
        yy = None if y is None else _as_f64(y)
        self._ck(self.lib.manisdp_set_dual(self._h, None if yy is None else _pf(yy), float(sigma)), "set_dual")

    def set_sigma(self, sigma):
        self._ck(self.lib.manisdp_set_sigma(self._h, float(sigma)), "set_sigma")

    def get_dual(self):
        y = np.empty(self.m)
        s = C.c_double()
        self._ck(self.lib.manisdp_get_dual(self._h, _pf(y), C.byref(s)), "get_dual")
        return y, s.value

    def slot_set(self, slot, A):
        A = _as_f64(A)
        self._ck(self.lib.manisdp_slot_set(self._h, slot, _pf(A), LAYOUT_ROWS), "slot_set")

    def slot_get(self, slot):
        out = np.empty((self.n_local, self.p))
        self._ck(self.lib.manisdp_slot_get(self._h, slot, _pf(out), LAYOUT_ROWS), "slot_get")
        return out

    # -- closures
    def cost(self):
        f = C.c_double()
        self._ck(self.lib.manisdp_cost(self._h, C.byref(f)), "cost")
        return f.value

    def grad(self):
        g = C.c_double()
        self._ck(self.lib.manisdp_grad(self._h, C.byref(g)), "grad")
        return self.slot_get(SLOT_G), g.value

    def hess(self, U):
        self.slot_set(SLOT_U, U)
        self._ck(self.lib.manisdp_hess(self._h), "hess")
        return self.slot_get(SLOT_H)

    def hess_bench(self, reps):
        ms = C.c_double()
        self._ck(self.lib.manisdp_hess_bench(self._h, reps, C.byref(ms)), "hess_bench")
        return ms.value

    def vec_bench(self, which, reps=20):
        """(ms per launch, algorithmic bytes) of a fused vector kernel: 0 retract, 1 project, 2 tCG update, 3 tCG dir"""
        ms, nb = C.c_double(), C.c_double()
        self._ck(self.lib.manisdp_vec_bench(self._h, which, reps, C.byref(ms), C.byref(nb)), "vec_bench")
        return ms.value, nb.value

    def retract(self, eta):
        self.slot_set(SLOT_U, eta)
        self._ck(self.lib.manisdp_retract(self._h, SLOT_U, SLOT_H), "retract")
        return self.slot_get(SLOT_H)

    def project(self, U):
        self.slot_set(SLOT_U, U)
        self._ck(self.lib.manisdp_project(self._h, SLOT_U, SLOT_H), "project")
        return self.slot_get(SLOT_H)

    # -- solver
    def tr_solve(self, maxiter=0, maxinner=0, tolgradnorm=0.0, use_graph=1, **kw):
        o = TrOptions()
        o.maxiter, o.maxinner, o.tolgradnorm, o.use_graph = int(maxiter), int(maxinner), float(tolgradnorm), int(use_graph)
        for k, v in kw.items():
            setattr(o, k, v)
        info = TrInfo()
        self._ck(self.lib.manisdp_tr_solve(self._h, C.byref(o), C.byref(info)), "tr_solve")
        return info

    def tr_log(self):
        cnt = C.c_int32()
        self._ck(self.lib.manisdp_tr_log(self._h, None, 0, C.byref(cnt)), "tr_log")
        buf = (TrIter * max(1, cnt.value))()
        self._ck(self.lib.manisdp_tr_log(self._h, buf, cnt.value, C.byref(cnt)), "tr_log")
        return [buf[i] for i in range(cnt.value)]

    # -- outer-loop pieces
    def kkt(self, delta=8, eig_tol=0.0, update_dual=1):
        k = KktInfo()
        self._ck(self.lib.manisdp_kkt(self._h, delta, eig_tol, update_dual, C.byref(k)), "kkt")
        return k

    def get_eigs(self, k, vectors=True):
        vals = np.empty(k)
        vecs = np.empty((self.n_local, k)) if vectors else None
        self._ck(self.lib.manisdp_get_eigs(self._h, _pf(vals), None if vecs is None else _pf(vecs), k), "get_eigs")
        return vals, vecs

    def rank_cut(self, theta, apply=True):
        r, pn = C.c_int64(), C.c_int64()
        self._ck(self.lib.manisdp_rank_cut(self._h, theta, int(apply), C.byref(r), C.byref(pn)), "rank_cut")
        return r.value, pn.value

    def escape(self, nne, alpha, line_search=0):
        self._ck(self.lib.manisdp_escape(self._h, nne, alpha, line_search), "escape")

    def line_search(self):
        a = C.c_double()
        self._ck(self.lib.manisdp_line_search(self._h, C.byref(a)), "line_search")
        return a.value

    def col_split(self):
        """column-sharded handle: keep this rank's ceil(p/world) columns of the factor (collective)"""
        self._ck(self.lib.manisdp_col_split(self._h), "col_split")

    def col_merge(self):
        """column-sharded handle: all-gather the column slices back into the full-width factor (collective)"""
        self._ck(self.lib.manisdp_col_merge(self._h), "col_merge")

    def index_split(self):
        """(i, j) int64 arrays of every stored entry of At, CSC order, as the device kernels index with them"""
        cnt = C.c_int64()
        self._ck(self.lib.manisdp_get_index_split(self._h, None, None, 0, C.byref(cnt)), "get_index_split")
        i = np.empty(cnt.value, dtype=np.int64)
        j = np.empty(cnt.value, dtype=np.int64)
        p64 = C.POINTER(C.c_int64)
        self._ck(self.lib.manisdp_get_index_split(self._h, i.ctypes.data_as(p64), j.ctypes.data_as(p64), cnt.value,
                                                  C.byref(cnt)), "get_index_split")
        return i, j

    def stats(self):
        s = Stats()
        self._ck(self.lib.manisdp_get_stats(self._h, C.byref(s)), "get_stats")
        return s

    # -- dual handles: the ADMM multipliers x (n x n) and w (K.f)
    def dual_state(self):
        x = np.empty(self.n * self.n)
        w = np.empty(max(1, self.nfree))
        self._ck(self.lib.manisdp_dual_get_state(self._h, _pf(x), _pf(w)), "dual_get_state")
        return x, w[:self.nfree]

    def dual_set_state(self, x=None, w=None):
        xx = None if x is None else _as_f64(np.asarray(x).ravel())
        ww = None if w is None or self.nfree == 0 else _as_f64(np.asarray(w).ravel())
        self._ck(self.lib.manisdp_dual_set_state(self._h, None if xx is None else _pf(xx),
                                                 None if ww is None else _pf(ww)), "dual_set_state")

    # -- multi-block handles (manisdp_mb_*): a point is a list of (n_i, p_i) arrays, one row per vertex of the block
    def _i64(self, v):
        a = np.ascontiguousarray(v, dtype=np.int64)
        assert a.shape == (len(self.block_sizes),)
        return a, a.ctypes.data_as(C.POINTER(C.c_int64))

    def mb_set_Y(self, blocks):
        assert len(blocks) == len(self.block_sizes)
        p, pp = self._i64([np.asarray(B).shape[1] for B in blocks])
        cat = np.concatenate([_as_f64(B).ravel() for B in blocks])
        self._ck(self.lib.manisdp_mb_set_Y(self._h, _pf(cat), pp), "mb_set_Y")

    def mb_widths(self):
        p, pp = self._i64(np.zeros(len(self.block_sizes)))
        self._ck(self.lib.manisdp_mb_get_widths(self._h, pp), "mb_get_widths")
        return [int(v) for v in p]

    def mb_get_Y(self):
        p = self.mb_widths()
        cat = np.empty(int(sum(n * q for n, q in zip(self.block_sizes, p))))
        self._ck(self.lib.manisdp_mb_get_Y(self._h, _pf(cat)), "mb_get_Y")
        out, o = [], 0
        for n, q in zip(self.block_sizes, p):
            out.append(cat[o:o + n * q].reshape(n, q).copy())
            o += n * q
        return out

    def mb_rand_Y(self, p, seed=0):
        _, pp = self._i64(p)
        self._ck(self.lib.manisdp_mb_rand_Y(self._h, pp, seed), "mb_rand_Y")

    def mb_split(self, A):
        """(N, pmax) slot array -> list of (n_i, p_i) blocks at the current widths"""
        out, r = [], 0
        for n, q in zip(self.block_sizes, self.mb_widths()):
            out.append(np.ascontiguousarray(A[r:r + n, :q]))
            r += n
        return out

    def mb_join(self, blocks):
        """list of (n_i, p_i) blocks -> (N, pmax) slot array (zero padded)"""
        pmax = self.p
        A = np.zeros((self.n, pmax))
        r = 0
        for B in blocks:
            A[r:r + B.shape[0], :B.shape[1]] = B
            r += B.shape[0]
        return A

    def mb_kkt(self, update_dual=1):
        k = KktInfo()
        t = len(self.block_sizes)
        dinfs = np.empty(t)
        nneg = np.empty(t, dtype=np.int32)
        self._ck(self.lib.manisdp_mb_kkt(self._h, int(update_dual), C.byref(k), _pf(dinfs),
                                         nneg.ctypes.data_as(C.POINTER(C.c_int32))), "mb_kkt")
        return k, dinfs, nneg

    def mb_block_eigs(self, blk, vectors=True):
        n = self.block_sizes[blk]
        vals = np.empty(n)
        vecs = np.empty((n, n)) if vectors else None
        self._ck(self.lib.manisdp_mb_get_block_eigs(self._h, blk, _pf(vals), None if vecs is None else _pf(vecs)),
                 "mb_get_block_eigs")
        return vals, vecs

    def mb_update(self, theta, delta, alpha, line_search=0, min_facsize=2):
        p, pp = self._i64(np.zeros(len(self.block_sizes)))
        self._ck(self.lib.manisdp_mb_update(self._h, float(theta), int(delta), float(alpha), int(line_search),
                                            int(min_facsize), pp), "mb_update")
        return [int(v) for v in p]


def nccl_unique_id() -> bytes:
    buf = (C.c_char * 128)()
    rc = load().manisdp_nccl_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise EngineError(f"manisdp_nccl_unique_id failed ({rc})")
    return bytes(buf)


class GroupHandle:
    """Single-process multi-GPU ONLYUNITDIAG solve (manisdp_group_*, csrc/group.cu): one caller thread, one worker
    thread + one column-sharded handle per device inside the library.  Same method names as Handle for what the
    ManiSDP_onlyunitdiag driver needs; the factor is merged (full width) between calls."""

    def __init__(self, n, C_csc, devices):
        import scipy.sparse as sp

        self.lib = load()
        self._g = _H()
        pb = Problem()
        pb.kind = ONLYUNITDIAG
        pb.n = n
        Cm = sp.csc_matrix(C_csc)
        Cm.sort_indices()
        jc, ir, pr = _as_u64(Cm.indptr), _as_u64(Cm.indices), _as_f64(Cm.data)
        pb.C_jc, pb.C_ir, pb.C_pr = _pu(jc), _pu(ir), _pf(pr)
        dev = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        rc = self.lib.manisdp_group_create(C.byref(self._g), C.byref(pb), len(devices), dev)
        if rc != 0:
            msg = self.lib.manisdp_group_last_error(None)
            raise EngineError(f"manisdp_group_create failed ({rc}): {msg.decode() if msg else ''}")
        self.n = self.n_local = n
        self.devices = list(devices)

    def _ck(self, rc, what):
        if rc != 0:
            msg = self.lib.manisdp_group_last_error(self._g)
            raise EngineError(f"group {what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if self._g:
            self.lib.manisdp_group_destroy(self._g)
            self._g = _H()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        s = Stats()
        self._ck(self.lib.manisdp_group_get_stats(self._g, C.byref(s)), "get_stats")
        return s

    @property
    def p(self):
        return int(self.stats().p)

    def set_Y(self, Y):
        Y = _as_f64(Y)
        assert Y.ndim == 2 and Y.shape[0] == self.n
        self._ck(self.lib.manisdp_group_set_Y(self._g, _pf(Y), Y.shape[1], LAYOUT_ROWS), "set_Y")

    def get_Y(self):
        out = np.empty((self.n, self.p))
        self._ck(self.lib.manisdp_group_get_Y(self._g, _pf(out), LAYOUT_ROWS), "get_Y")
        return out

    def rand_Y(self, p, seed=0):
        self._ck(self.lib.manisdp_group_rand_Y(self._g, p, seed), "rand_Y")

    def cost(self):
        f = C.c_double()
        self._ck(self.lib.manisdp_group_cost(self._g, C.cast(C.byref(f), _f64p)), "cost")
        return f.value

    def tr_solve(self, maxiter=0, maxinner=0, tolgradnorm=0.0, use_graph=1, **kw):
        o = TrOptions()
        o.maxiter, o.maxinner, o.tolgradnorm, o.use_graph = int(maxiter), int(maxinner), float(tolgradnorm), int(use_graph)
        for k, v in kw.items():
            setattr(o, k, v)
        info = TrInfo()
        self._ck(self.lib.manisdp_group_tr_solve(self._g, C.byref(o), C.byref(info)), "tr_solve")
        return info

    def kkt(self, delta=8, eig_tol=0.0, update_dual=0):
        k = KktInfo()
        self._ck(self.lib.manisdp_group_kkt(self._g, delta, eig_tol, update_dual, C.byref(k)), "kkt")
        return k

    def rank_cut(self, theta, apply=True):
        r, pn = C.c_int64(), C.c_int64()
        self._ck(self.lib.manisdp_group_rank_cut(self._g, theta, int(apply), C.byref(r), C.byref(pn)), "rank_cut")
        return r.value, pn.value

    def escape(self, nne, alpha, line_search=0):
        self._ck(self.lib.manisdp_group_escape(self._g, nne, alpha, line_search), "escape")

    def line_search(self):
        a = C.c_double()
        self._ck(self.lib.manisdp_group_line_search(self._g, C.cast(C.byref(a), _f64p)), "line_search")
        return a.value
