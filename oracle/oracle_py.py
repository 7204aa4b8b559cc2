"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs -- never by the product package ``permon_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DECIDE = -1.0
INFINITY = 1.7976931348623157e308 / 4.0
NINFINITY = -INFINITY
EPS = 2.2204460492503131e-16

EXP = {"std": 0, "projcg": 1, "gf": 2, "g": 3, "gfgr": 4, "ggr": 5}
LEN = {"fixed": 0, "opt": 1, "optapprox": 2, "bb": 3}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class Op(C.Structure):
    _fields_ = [("kind", C.c_int), ("n", C.c_int), ("ia", _ip), ("ja", _ip), ("a", _dp), ("d", C.c_int),
                ("ia2", _ip), ("ja2", _ip), ("a2", _dp), ("twork", _dp), ("m", C.c_int), ("B", _dp),
                ("rho", C.c_double), ("bwork", _dp),
                ("pmode", C.c_int), ("pm", C.c_int), ("PG", _dp), ("porth", C.c_int), ("pw1", _dp), ("pw2", _dp), ("bimplicit", C.c_int)]


class Box(C.Structure):
    _fields_ = [("n", C.c_int), ("nis", C.c_int), ("is_", _ip), ("lb", _dp), ("ub", _dp), ("astol", C.c_double)]


class MpgpOpts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("divtol", C.c_double), ("max_it", C.c_int),
                ("alpha_user", C.c_double), ("alpha_direct", C.c_int), ("gamma", C.c_double),
                ("maxeig", C.c_double), ("maxeig_tol", C.c_double), ("maxeig_iter", C.c_int),
                ("exptype", C.c_int), ("explengthtype", C.c_int), ("resetalpha", C.c_int),
                ("fallback", C.c_int), ("fallback2", C.c_int), ("nthreads", C.c_int)]


class MpgpResult(C.Structure):
    _fields_ = [("its", C.c_int), ("reason", C.c_int), ("nmv", C.c_int), ("ncg", C.c_int), ("nexp", C.c_int),
                ("nprop", C.c_int), ("nfinc", C.c_int), ("nfall", C.c_int), ("rnorm", C.c_double),
                ("alpha", C.c_double), ("maxeig", C.c_double), ("maxeig_its", C.c_int),
                ("norm_rhs", C.c_double), ("ttol", C.c_double), ("seconds", C.c_double)]


class Trace(C.Structure):
    _fields_ = [("cap", C.c_int), ("len", C.c_int), ("step", C.c_char_p), ("rnorm", _dp), ("gfnorm", _dp),
                ("gcnorm", _dp), ("alpha", _dp)]


class SmalxeOpts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("divtol", C.c_double), ("max_it", C.c_int),
                ("M1_user", C.c_double), ("M1_direct", C.c_int), ("M1_update", C.c_double),
                ("rho_user", C.c_double), ("rho_direct", C.c_int), ("rho_update", C.c_double),
                ("rho_update_late", C.c_double), ("eta_user", C.c_double), ("eta_direct", C.c_int),
                ("rtol_E", C.c_double), ("update_threshold", C.c_double), ("maxeig", C.c_double),
                ("maxeig_tol", C.c_double), ("maxeig_iter", C.c_int), ("inject_maxeig", C.c_int),
                ("inject_maxeig_set", C.c_int), ("inner_iter_min", C.c_int), ("inner_no_gtol_stop", C.c_int),
                ("knoll", C.c_int), ("get_lambda", C.c_int), ("inner", MpgpOpts),
                ("implicit_orth", C.c_int), ("lag_enabled", C.c_int), ("lag_offset", C.c_int), ("Jstart", C.c_int), ("Jstep", C.c_int),
                ("Jend", C.c_int), ("lag_lower", C.c_double), ("lag_upper", C.c_double)]


class SmalxeResult(C.Structure):
    _fields_ = [("outer_its", C.c_int), ("reason", C.c_int), ("inner_its_accu", C.c_int),
                ("inner_reason_last", C.c_int), ("M1_updates", C.c_int), ("M1_hits", C.c_int),
                ("eta_hits", C.c_int), ("rho_updates", C.c_int), ("state", C.c_int), ("nmv", C.c_int),
                ("ncg", C.c_int), ("nexp", C.c_int), ("nprop", C.c_int), ("rnorm", C.c_double),
                ("normBu", C.c_double), ("M1", C.c_double), ("rho", C.c_double), ("maxeig", C.c_double),
                ("maxeig_inner", C.c_double), ("alpha_inner", C.c_double), ("eta", C.c_double),
                ("seconds", C.c_double)]


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (gcc, seconds)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("permon_oracle.c", "permon_oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "CC": "/usr/bin/gcc"})
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_dot.restype = C.c_double
        L.orc_norm2.restype = C.c_double
        L.orc_qpc_feas.restype = C.c_double
        L.orc_objective.restype = C.c_double
        L.orc_max_eigenvalue.restype = C.c_int
        L.orc_get_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


class Operator:
    """Holds the numpy arrays alive next to the C struct."""

    def __init__(self, ia, ja, a, n=None, second=None):
        self.ia, self.ja, self.a = i32(ia), i32(ja), f64(a)
        self.n = int(n if n is not None else len(self.ia) - 1)
        self.c = Op()
        self.c.kind = 0
        self.c.n = self.n
        self.c.ia, self.c.ja, self.c.a = _i(self.ia), _i(self.ja), _d(self.a)
        if second is not None:  # A = M1 * M2; (ia,ja,a) is M1 (n x d), second is M2 (d x n)
            self.ia2, self.ja2, self.a2 = i32(second[0]), i32(second[1]), f64(second[2])
            self.d = len(self.ia2) - 1
            self.tw = np.zeros(self.d)
            self.c.kind = 1
            self.c.d = self.d
            self.c.ia2, self.c.ja2, self.c.a2, self.c.twork = _i(self.ia2), _i(self.ja2), _d(self.a2), _d(self.tw)
        self.c.m = 0

    def set_projector(self, G, mode=2):
        """A -> P A P (mode 2) or P A (mode 1), P = I - G'(G G')^-1 G  (QPTEnforceEqByProjector)"""
        self.PG = f64(G).reshape(-1, self.n)
        self.pw = np.zeros((2, self.n))
        self.c.pmode = int(mode)
        self.c.pm = self.PG.shape[0]
        self.c.PG = _d(self.PG)
        self.c.porth = int(lib().orc_rows_orthonormal(C.c_int(self.n), C.c_int(self.c.pm), _d(self.PG)))
        self.c.pw1, self.c.pw2 = _d(self.pw[0]), _d(self.pw[1])

    def set_penalty(self, B, rho):
        self.B = f64(B).reshape(-1, self.n)
        self.bw = np.zeros(self.B.shape[0])
        self.c.m = self.B.shape[0]
        self.c.B = _d(self.B)
        self.c.rho = float(rho)
        self.c.bwork = _d(self.bw)

    def apply(self, x):
        x = f64(x)
        y = np.empty(self.n)
        lib().orc_op_apply(C.byref(self.c), _d(x), _d(y))
        return y


class BoxC:
    def __init__(self, n, lb=None, ub=None, is_=None):
        self.lb, self.ub, self.is_ = f64(lb), f64(ub), i32(is_)
        self.c = Box()
        self.c.n = int(n)
        self.c.nis = -1 if is_ is None else len(self.is_)
        self.c.is_ = _i(self.is_)
        self.c.lb, self.c.ub = _d(self.lb), _d(self.ub)
        self.c.astol = 10 * EPS
        self.nsub = int(n) if is_ is None else len(self.is_)


def mpgp_opts(**kw) -> MpgpOpts:
    o = MpgpOpts()
    lib().orc_default_mpgp_opts(C.byref(o))
    for k, v in kw.items():
        if k == "exptype" and isinstance(v, str):
            v = EXP[v]
        if k == "explengthtype" and isinstance(v, str):
            v = LEN[v]
        if not hasattr(o, k):
            raise KeyError(k)
        setattr(o, k, v)
    return o


def smalxe_opts(inner=None, **kw) -> SmalxeOpts:
    o = SmalxeOpts()
    lib().orc_default_smalxe_opts(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise KeyError(k)
        setattr(o, k, v)
    if inner:
        for k, v in inner.items():
            if k == "exptype" and isinstance(v, str):
                v = EXP[v]
            if k == "explengthtype" and isinstance(v, str):
                v = LEN[v]
            setattr(o.inner, k, v)
    return o


def mpgp_solve(op: Operator, b, box: BoxC, x0=None, opts: MpgpOpts | None = None, trace_cap: int = 0):
    b = f64(b)
    x = np.zeros(op.n) if x0 is None else f64(x0).copy()
    opts = opts or mpgp_opts()
    res = MpgpResult()
    tr = None
    if trace_cap:
        tr = Trace()
        bufs = dict(step=C.create_string_buffer(trace_cap + 1), rnorm=np.zeros(trace_cap), gfnorm=np.zeros(trace_cap),
                    gcnorm=np.zeros(trace_cap), alpha=np.zeros(trace_cap))
        tr.cap = trace_cap
        tr.step = C.cast(bufs["step"], C.c_char_p)
        tr.rnorm, tr.gfnorm, tr.gcnorm, tr.alpha = (_d(bufs[k]) for k in ("rnorm", "gfnorm", "gcnorm", "alpha"))
    lib().orc_mpgp_solve(C.byref(op.c), _d(b), C.byref(box.c), _d(x), C.byref(opts), C.byref(res),
                         C.byref(tr) if tr is not None else None)
    out = {f[0]: getattr(res, f[0]) for f in MpgpResult._fields_}
    if tr is not None:
        k = tr.len
        out["trace"] = dict(step=bufs["step"].raw[:k].decode(), rnorm=bufs["rnorm"][:k].copy(),
                            gfnorm=bufs["gfnorm"][:k].copy(), gcnorm=bufs["gcnorm"][:k].copy(),
                            alpha=bufs["alpha"][:k].copy())
    return x, out


def smalxe_solve(op: Operator, b, box: BoxC, B, c=None, x0=None, opts: SmalxeOpts | None = None):
    b = f64(b)
    Bm = f64(B).reshape(-1, op.n)
    m = Bm.shape[0]
    cc = f64(c)
    x = np.zeros(op.n) if x0 is None else f64(x0).copy()
    opts = opts or smalxe_opts()
    res = SmalxeResult()
    Btl = np.zeros(op.n)
    lam = np.zeros(m)
    lib().orc_smalxe_solve(C.byref(op.c), _d(b), C.byref(box.c), C.c_int(m), _d(Bm), _d(cc), _d(x), C.byref(opts),
                           C.byref(res), _d(Btl), _d(lam))
    out = {f[0]: getattr(res, f[0]) for f in SmalxeResult._fields_}
    out["Bt_lambda"] = Btl
    out["lambda"] = lam
    return x, out


class LinOpts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("divtol", C.c_double), ("max_it", C.c_int), ("nthreads", C.c_int)]


class LinResult(C.Structure):
    _fields_ = [("its", C.c_int), ("reason", C.c_int), ("rnorm", C.c_double), ("norm_rhs", C.c_double), ("seconds", C.c_double)]


def lin_opts(**kw) -> LinOpts:
    o = LinOpts()
    lib().orc_default_lin_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def cg_solve(op: Operator, b, x0=None, opts: LinOpts | None = None):
    """QPSKSP = KSPCG without preconditioner (restated; parity unpinned, see permon_oracle.c)"""
    b = f64(b)
    x = np.zeros(op.n) if x0 is None else f64(x0).copy()
    res = LinResult()
    lib().orc_cg_solve(C.byref(op.c), _d(b), _d(x), C.byref(opts or lin_opts()), C.byref(res))
    return x, {f[0]: getattr(res, f[0]) for f in LinResult._fields_}


def pcpg_solve(op: Operator, b, G, c=None, x0=None, opts: LinOpts | None = None):
    """QPSPCPG (pcpg.c:49-131) for min 1/2 x'Ax - b'x s.t. G x = c"""
    b = f64(b)
    Gm = f64(G).reshape(-1, op.n)
    x = np.zeros(op.n) if x0 is None else f64(x0).copy()
    res = LinResult()
    cc = None if c is None else f64(c)
    lib().orc_pcpg_solve(C.byref(op.c), _d(b), C.c_int(Gm.shape[0]), _d(Gm), _d(cc) if cc is not None else None, _d(x), C.byref(opts or lin_opts()),
                         C.byref(res))
    return x, {f[0]: getattr(res, f[0]) for f in LinResult._fields_}


def orth_rows(B, c=None, kind="gs"):
    """QPTOrthonormalizeEq's MatOrthRows, explicit form: -> (T B, T c, T)"""
    Bm = f64(B)
    m, n = Bm.shape
    TB, T = np.empty((m, n)), np.empty((m, m))
    cc = None if c is None else f64(c)
    Tc = None if c is None else np.empty(m)
    rc = lib().orc_orth_rows(C.c_int(n), C.c_int(m), _d(Bm), _d(cc) if cc is not None else None, C.c_int({"gs": 1, "cholesky": 3}[kind]), _d(TB),
                             _d(Tc) if Tc is not None else None, _d(T))
    if rc:
        raise RuntimeError(f"orc_orth_rows failed ({rc})")
    return TB, Tc, T


def apply_P(G, v):
    Gm = f64(G)
    m, n = Gm.shape
    out = np.empty(n)
    lib().orc_apply_P(C.c_int(n), C.c_int(m), _d(Gm), _d(f64(v)), _d(out))
    return out


def homogenize(op: Operator, b, box, G, c):
    """QPTHomogenizeEq: -> (xtilde, b_h, lb_h, ub_h)"""
    Gm = f64(G)
    m, n = Gm.shape
    xt, bh = np.empty(n), np.empty(n)
    lbh = None if box is None or box.lb is None else np.empty(len(box.lb))
    ubh = None if box is None or box.ub is None else np.empty(len(box.ub))
    lib().orc_homogenize(C.byref(op.c), _d(f64(b)), C.byref(box.c) if box is not None else None, C.c_int(m), _d(Gm), _d(f64(c)), _d(xt), _d(bh),
                         _d(lbh) if lbh is not None else None, _d(ubh) if ubh is not None else None)
    return xt, bh, lbh, ubh


def max_eigenvalue(op: Operator, tol=DECIDE, maxits=-1):
    lam = C.c_double()
    its = lib().orc_max_eigenvalue(C.byref(op.c), C.c_double(tol), C.c_int(maxits), C.byref(lam))
    return lam.value, its


def objective(op: Operator, b, x):
    b, x = f64(b), f64(x)
    return lib().orc_objective(C.byref(op.c), _d(b), _d(x))


def box_multipliers(op: Operator, b, box: BoxC, x, Bt_lambda=None):
    b, x, btl = f64(b), f64(x), f64(Bt_lambda)
    llb = np.zeros(box.nsub) if box.lb is not None else None
    lub = np.zeros(box.nsub) if box.ub is not None else None
    lib().orc_box_multipliers(C.byref(op.c), _d(b), _d(btl), C.byref(box.c), _d(x), _d(llb), _d(lub))
    return llb, lub


def kkt(op: Operator, b, box: BoxC, x, llb, lub, Bt_lambda=None):
    b, x, btl = f64(b), f64(x), f64(Bt_lambda)
    out = np.zeros(8)
    lib().orc_kkt(C.byref(op.c), _d(b), _d(btl), C.byref(box.c), _d(x), _d(f64(llb)), _d(f64(lub)), _d(out))
    return out


def qpc_project(box: BoxC, x):
    x = f64(x)
    y = np.empty_like(x)
    lib().orc_qpc_project(C.byref(box.c), _d(x), _d(y))
    return y


def qpc_grads(box: BoxC, x, g):
    x, g = f64(x), f64(g)
    gf, gc = np.empty_like(x), np.empty_like(x)
    lib().orc_qpc_grads(C.byref(box.c), _d(x), _d(g), _d(gf), _d(gc))
    return gf, gc


def qpc_gradreduced(box: BoxC, x, gf, alpha):
    x, gf = f64(x), f64(gf)
    gr = np.empty_like(x)
    lib().orc_qpc_gradreduced(C.byref(box.c), _d(x), _d(gf), C.c_double(alpha), _d(gr))
    return gr


def qpc_feas(box: BoxC, x, d):
    x, d = f64(x), f64(d)
    return lib().orc_qpc_feas(C.byref(box.c), _d(x), _d(d))
