"""CPU: the MATLAB MEX gateway source compiles against include/manisdp_b200.h (syntax + type check with a stub mex.h;
MATLAB itself is absent from the build container) and only calls symbols the header declares."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gateway_compiles_against_header():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tests", "mex_stub"), os.path.join(ROOT, "matlab", "manisdp_mex.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_gateway_uses_only_declared_entry_points():
    hdr = open(os.path.join(ROOT, "include", "manisdp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(manisdp_[a-z_A-Z0-9]+)\s*\(", hdr))
    src = open(os.path.join(ROOT, "matlab", "manisdp_mex.cpp")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    used = set(re.findall(r"\b(manisdp_[a-z_A-Z0-9]+)\s*\(", src)) - {"manisdp_mex"}
    assert used <= declared, used - declared


def test_matlab_drivers_keep_reference_signatures():
    sigs = {"ManiSDP_onlyunitdiag.m": "function [X, obj, data] = ManiSDP_onlyunitdiag(C, options)",
            "ManiSDP_unitdiag.m": "function [X, obj, data] = ManiSDP_unitdiag(At, b, c, K, options)",
            "ManiSDP_unittrace.m": "function [X, obj, data] = ManiSDP_unittrace(At, b, c, K, options)",
            "ManiSDP.m": "function [X, obj, data] = ManiSDP(At, b, c, K, options)",
            "ManiSDP_multiblock.m": "function [X, obj, data] = ManiSDP_multiblock(At, b, c, K, options)",
            "ManiDSDP_unitdiag.m": "function [X, obj, data] = ManiDSDP_unitdiag(A, b, c, K, options)"}
    for f, s in sigs.items():
        assert open(os.path.join(ROOT, "matlab", f)).readline().strip() == s


# ---- executing the gateway (functional mex.h stand-in, tests/mex_stub/mex_stub.cpp) ------------------------------------
import ctypes as C  # noqa: E402

import numpy as np  # noqa: E402
import pytest  # noqa: E402


class Mex:
    """drives mexFunction of the unmodified matlab/manisdp_mex.cpp through the stub's C harness"""

    def __init__(self):
        import sys
        sys.path.insert(0, os.path.join(ROOT, "tests", "mex_stub"))
        import build_harness
        self.lib = C.CDLL(build_harness.build())
        L = self.lib
        vp = C.c_void_p
        L.stub_string.restype = vp
        L.stub_string.argtypes = [C.c_char_p]
        L.stub_dense.restype = vp
        L.stub_dense.argtypes = [C.c_size_t, C.c_size_t, C.POINTER(C.c_double)]
        L.stub_sparse.restype = vp
        L.stub_sparse.argtypes = [C.c_size_t, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_double)]
        L.stub_struct.restype = vp
        L.stub_struct.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double)]
        L.stub_free.argtypes = [vp]
        L.stub_rows.restype = L.stub_cols.restype = C.c_size_t
        L.stub_rows.argtypes = L.stub_cols.argtypes = [vp]
        L.stub_kind.argtypes = [vp]
        L.stub_data.restype = C.POINTER(C.c_double)
        L.stub_data.argtypes = [vp]
        L.stub_u64.restype = C.c_uint64
        L.stub_u64.argtypes = [vp]
        L.stub_field.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double)]
        L.stub_call.argtypes = [C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_char_p, C.c_char_p, C.c_size_t]
        L.stub_cell.restype = vp
        L.stub_cell.argtypes = [C.c_size_t]
        L.stub_cell_set.argtypes = [vp, C.c_size_t, vp]
        L.stub_cell_get.restype = vp
        L.stub_cell_get.argtypes = [vp, C.c_size_t]

    def arr(self, x):
        import scipy.sparse as sp
        L = self.lib
        if isinstance(x, str):
            return L.stub_string(x.encode())
        if isinstance(x, list):  # cell array (column) of matrices
            cell = L.stub_cell(len(x))
            for i, e in enumerate(x):
                L.stub_cell_set(cell, i, self.arr(e))
            return cell
        if isinstance(x, dict):
            names = (C.c_char_p * len(x))(*[k.encode() for k in x])
            vals = (C.c_double * len(x))(*[float(v) for v in x.values()])
            return L.stub_struct(len(x), names, vals)
        if sp.issparse(x):
            m = sp.csc_matrix(x)
            m.sort_indices()
            jc = np.ascontiguousarray(m.indptr, dtype=np.uint64)
            ir = np.ascontiguousarray(m.indices, dtype=np.uint64)
            pr = np.ascontiguousarray(m.data, dtype=np.float64)
            return L.stub_sparse(m.shape[0], m.shape[1], jc.ctypes.data_as(C.POINTER(C.c_uint64)),
                                 ir.ctypes.data_as(C.POINTER(C.c_uint64)), pr.ctypes.data_as(C.POINTER(C.c_double)))
        a = np.atleast_2d(np.asarray(x, dtype=np.float64))
        f = np.asfortranarray(a)
        return L.stub_dense(a.shape[0], a.shape[1], f.ctypes.data_as(C.POINTER(C.c_double)))

    def call(self, nlhs, *args, handle=None):
        """returns the list of outputs (numpy arrays, dicts are read with .field) or raises RuntimeError(id: msg)"""
        L = self.lib
        ins = []
        for i, a in enumerate(args):
            ins.append(handle if (i == 1 and handle is not None and a is None) else self.arr(a))
        prhs = (C.c_void_p * len(ins))(*ins)
        plhs = (C.c_void_p * max(1, nlhs))()
        eid, msg = C.create_string_buffer(256), C.create_string_buffer(2048)
        rc = L.stub_call(nlhs, plhs, len(ins), prhs, eid, msg, 2048)
        for i, a in enumerate(ins):
            if not (i == 1 and handle is not None and args[i] is None):
                L.stub_free(a)
        if rc != 0:
            raise RuntimeError(f"{eid.value.decode()}: {msg.value.decode()}")
        return [plhs[i] for i in range(nlhs)]

    def mat(self, a):
        m, n = self.lib.stub_rows(a), self.lib.stub_cols(a)
        out = np.ctypeslib.as_array(self.lib.stub_data(a), shape=(n, m)).T.copy()  # column-major -> (m, n)
        self.lib.stub_free(a)
        return out

    def cell(self, a, count):
        """cell array of matrices -> list of (m, n) arrays"""
        out = []
        for i in range(count):
            e = self.lib.stub_cell_get(a, i)
            m, n = self.lib.stub_rows(e), self.lib.stub_cols(e)
            out.append(np.ctypeslib.as_array(self.lib.stub_data(e), shape=(n, m)).T.copy())
        self.lib.stub_free(a)
        return out

    def field(self, s, name):
        v = C.c_double()
        assert self.lib.stub_field(s, name.encode(), C.byref(v)) == 0, name
        return v.value


def test_gateway_links_and_raises_matlab_style_errors():
    """CPU: the harness library (gateway + stub + libmanisdp_b200.so) links and loads, and argument errors come back
    as mexErrMsgIdAndTxt identifiers (reference convention src/C-files/innerc.cpp:5-10) -- no GPU call is made."""
    mex = Mex()
    with pytest.raises(RuntimeError, match="ManiSDP:b200:arg: first argument must be a command string"):
        mex.call(0, 1.0)
    with pytest.raises(RuntimeError, match="ManiSDP:b200:arg: bad handle"):
        mex.call(0, "cost", 3.0)


@pytest.mark.gpu
def test_gateway_executes_the_hot_path_like_the_ctypes_binding():
    """GPU: create -> set_Y -> cost -> tr_solve -> kkt -> rank_cut -> get_Y -> destroy through mexFunction, on G11 with
    the p x n layout ManiSDP_onlyunitdiag.m uses; every number equals the ctypes binding's on the same inputs (the two
    bindings sit on the same C ABI)."""
    from manisdp_matlab_b200 import Handle, problems as P
    d = np.load(os.path.join(ROOT, "tests", "golden", "G11.npz"))
    Cm = P.maxcut_C(int(d["n"]), d["ei"].astype(np.int64), d["ej"].astype(np.int64), d["w"].astype(np.float64))
    n, p = Cm.shape[0], 12
    rng = np.random.default_rng(3)
    Y0 = rng.standard_normal((n, p))
    Y0 /= np.linalg.norm(Y0, axis=1, keepdims=True)
    with Handle("onlyunitdiag", n, C_csc=Cm) as h:
        h.set_Y(Y0)
        f_ref = h.cost()
        info = h.tr_solve(maxiter=8, maxinner=30, tolgradnorm=1e-8, use_graph=1)
        k = h.kkt(8, 1e-9, 0)
        r_ref, _ = h.rank_cut(1e-1, apply=False)
        Y_ref = h.get_Y()
    mex = Mex()
    (hd,) = mex.call(1, "create", 0.0, float(n), Cm)
    assert mex.lib.stub_kind(hd) == 4 and mex.lib.stub_is_locked() == 1  # uint64 handle, mexLock taken
    mex.call(0, "set_Y", None, Y0.T, 0.0, handle=hd)  # p x n, layout 0
    (f,) = mex.call(1, "cost", None, handle=hd)
    assert mex.mat(f)[0, 0] == f_ref
    (s,) = mex.call(1, "tr_solve", None, dict(maxiter=8, maxinner=30, tolgradnorm=1e-8, use_graph=1), handle=hd)
    assert mex.field(s, "cost") == info.cost and mex.field(s, "hv_count") == info.hv_count
    assert mex.field(s, "gradnorm") == info.gradnorm and mex.field(s, "iters") == info.iters
    (kk,) = mex.call(1, "kkt", None, 8.0, 1e-9, 0.0, handle=hd)
    assert mex.field(kk, "obj") == k.obj
    assert abs(mex.field(kk, "dinf") - k.dinf) <= 1e-9 * (1 + abs(k.dinf)) and mex.field(kk, "nneg") == k.nneg
    r, pp = mex.call(2, "rank_cut", None, 1e-1, 0.0, handle=hd)
    assert mex.mat(r)[0, 0] == r_ref and mex.mat(pp)[0, 0] == p
    (Y,) = mex.call(1, "get_Y", None, 0.0, handle=hd)
    Ym = mex.mat(Y)
    assert Ym.shape == (p, n) and np.array_equal(Ym.T, Y_ref)
    with pytest.raises(RuntimeError, match="ManiSDP:b200:arg: unknown command"):
        mex.call(0, "no_such_command", None, handle=hd)
    mex.call(0, "destroy", None, handle=hd)
    with pytest.raises(RuntimeError, match="stale handle"):
        mex.call(0, "cost", None, handle=hd)
    mex.lib.stub_free(hd)
    mex.lib.stub_run_at_exit()


@pytest.mark.gpu
def test_gateway_drives_a_multiblock_handle():
    """GPU: the multi-block commands of the gateway (cell arrays of p_i x n_i factors, K.s / K.nob at create) give the
    numbers of the ctypes binding: create -> mb_set_Y -> tr_solve -> mb_kkt -> mb_update -> mb_get_Y."""
    from instances import generators as G
    from manisdp_matlab_b200 import Handle
    At, b, c, K = G.multiblock_random([8, 6, 7, 2], 2, 5, 4)
    ns, nob = K["s"], K["nob"]
    rng = np.random.default_rng(2)
    Y0 = []
    for i, n in enumerate(ns):
        B = rng.standard_normal((n, 2))
        if i < nob:
            B /= np.linalg.norm(B, axis=1, keepdims=True)
        Y0.append(B)
    m = At.shape[1]
    with Handle("multiblock", sum(ns), At=At, b=b, c=c, block_sizes=ns, nob=nob) as h:
        h.set_dual(np.zeros(m), 0.1)
        h.mb_set_Y(Y0)
        info = h.tr_solve(maxiter=4, maxinner=20, tolgradnorm=1e-8, use_graph=1)
        k, dinfs, _ = h.mb_kkt(1)
        pn = h.mb_update(1e-2, 8, 0.1, 0, 2)
        Yref = h.mb_get_Y()
    mex = Mex()
    nsd = np.asarray(ns, dtype=np.float64).reshape(1, -1)
    (hd,) = mex.call(1, "create", 4.0, float(sum(ns)), At, b.reshape(-1, 1), c.reshape(-1, 1), nsd, float(nob))
    mex.call(0, "set_dual", None, np.zeros((m, 1)), 0.1, handle=hd)
    mex.call(0, "mb_set_Y", None, [B.T for B in Y0], handle=hd)  # Y{i} is p_i x n_i in the reference
    (s,) = mex.call(1, "tr_solve", None, dict(maxiter=4, maxinner=20, tolgradnorm=1e-8, use_graph=1), handle=hd)
    assert mex.field(s, "cost") == info.cost and mex.field(s, "hv_count") == info.hv_count
    kk, dd = mex.call(2, "mb_kkt", None, 1.0, nsd, handle=hd)
    assert mex.field(kk, "obj") == k.obj and mex.field(kk, "pinf") == k.pinf and mex.field(kk, "dinf") == k.dinf
    assert np.array_equal(mex.mat(dd).ravel(), dinfs)
    (pp,) = mex.call(1, "mb_update", None, 1e-2, 8.0, 0.1, 0.0, 2.0, nsd, handle=hd)
    assert [int(v) for v in mex.mat(pp).ravel()] == pn
    (Yc,) = mex.call(1, "mb_get_Y", None, nsd, handle=hd)
    for Ym, Yr in zip(mex.cell(Yc, len(ns)), Yref):
        assert np.array_equal(Ym.T, Yr)
    mex.call(0, "destroy", None, handle=hd)
    mex.lib.stub_free(hd)
    mex.lib.stub_run_at_exit()


@pytest.mark.gpu
def test_gateway_drives_a_dual_handle():
    """GPU: a MANISDP_DUAL_UNITDIAG handle through mexFunction (create with At = A(:,K.f+1:end)', dAAt, B, cf as
    matlab/ManiDSDP_unitdiag.m passes them) gives the numbers of the ctypes binding: tr_solve, the ADMM step, y, x, w."""
    import scipy.sparse as sp
    from instances import generators as G
    from manisdp_matlab_b200 import Handle
    d = np.load(os.path.join(ROOT, "tests", "golden", "bqp_10_1.npz"))
    A, b, dAAt, mb = G.bqpsos(d["Q"], d["e"], 10)
    b = b / np.abs(b).max()
    m = A.shape[0]
    B = sp.csc_matrix((np.ones(1), (np.zeros(1, dtype=int), np.zeros(1, dtype=int))), shape=(m, 1))
    cf, cp = np.array([1.0]), np.zeros(mb * mb)
    rng = np.random.default_rng(6)
    Y0 = rng.standard_normal((mb, 5))
    Y0 /= np.linalg.norm(Y0, axis=1, keepdims=True)
    with Handle("dual_unitdiag", mb, At=A.T.tocsc(), b=b, c=cp, dAAt=dAAt, B=B, cf=cf) as h:
        h.set_sigma(1e-3)
        h.set_Y(Y0)
        info = h.tr_solve(maxiter=4, maxinner=20, tolgradnorm=1e-8, use_graph=1)
        k = h.kkt(8, 1e-9, 1)
        y_ref, _ = h.get_dual()
        x_ref, w_ref = h.dual_state()
    mex = Mex()
    (hd,) = mex.call(1, "create", 5.0, float(mb), A.T.tocsc(), b.reshape(-1, 1), cp.reshape(-1, 1), dAAt.reshape(-1, 1), B,
                     cf.reshape(-1, 1))
    mex.call(0, "set_sigma", None, 1e-3, handle=hd)
    mex.call(0, "set_Y", None, Y0.T, 0.0, handle=hd)
    (s,) = mex.call(1, "tr_solve", None, dict(maxiter=4, maxinner=20, tolgradnorm=1e-8, use_graph=1), handle=hd)
    assert mex.field(s, "cost") == info.cost and mex.field(s, "hv_count") == info.hv_count
    (kk,) = mex.call(1, "kkt", None, 8.0, 1e-9, 1.0, handle=hd)
    assert mex.field(kk, "obj") == k.obj and mex.field(kk, "pinf") == k.pinf and mex.field(kk, "gap") == k.gap
    (yy,) = mex.call(1, "get_dual", None, handle=hd)
    assert np.array_equal(mex.mat(yy).ravel(), y_ref)
    xx, ww = mex.call(2, "dual_state", None, 1.0, handle=hd)
    assert np.array_equal(mex.mat(xx).ravel(), x_ref) and np.array_equal(mex.mat(ww).ravel(), w_ref)
    mex.call(0, "destroy", None, handle=hd)
    mex.lib.stub_free(hd)
    mex.lib.stub_run_at_exit()
