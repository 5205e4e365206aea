"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/piso_oracle.c (the CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package ``fluidgym_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpiso_oracle.so")
SRC = os.path.join(HERE, "piso_oracle.c")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-ffp-contract=off",
                               "-o", LIB, SRC, "-lm"])
    return LIB


class StepCfg(C.Structure):
    _fields_ = [("corrector_steps", C.c_int), ("adv_nonortho_steps", C.c_int), ("p_nonortho_steps", C.c_int),
                ("nonortho", C.c_int), ("adv_tol", C.c_float), ("p_tol", C.c_float), ("maxit", C.c_int),
                ("bicg_iters", C.c_int * 2), ("cg_iters", C.c_int * 8), ("n_cg", C.c_int)]


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Oracle:
    """CPU oracle for one domain (one environment).  Fields are component-major global vectors:
    ``u[2, N]``, ``p[N]``, boundary velocities ``bvel[2, NB]`` -- the layout of the reference's global
    result vectors (K.cu:183-190)."""

    def __init__(self, sizes, btype, bconn, T, bT, bvel, visc, conn_corner_offset=1,
                 corrector_steps=2, adv_nonortho_steps=1, p_nonortho_steps=1, nonortho=True,
                 adv_tol=1e-5, p_tol=1e-5, maxit=5000):
        self.lib = C.CDLL(build())
        L = self.lib
        L.orc_create.restype = C.c_void_p
        L.orc_work_create.restype = C.c_void_p
        L.orc_work_field.restype = C.POINTER(C.c_float)
        L.orc_work_index.restype = C.POINTER(C.c_int)
        L.orc_bvel.restype = C.POINTER(C.c_float)
        L.orc_flux_balance.restype = C.c_float
        L.orc_max_velocity.restype = C.c_float
        sizes = np.ascontiguousarray(sizes, dtype=np.int32)
        btype = np.ascontiguousarray(btype, dtype=np.int32)
        bconn = np.ascontiguousarray(bconn, dtype=np.int32)
        T = np.ascontiguousarray(T, dtype=np.float32)
        bT = np.ascontiguousarray(bT, dtype=np.float32)
        bvel = np.ascontiguousarray(bvel, dtype=np.float32)
        self.dom = C.c_void_p(L.orc_create(len(sizes), _ip(sizes), _ip(btype), _ip(bconn), _fp(T), _fp(bT),
                                           _fp(bvel), C.c_float(visc), conn_corner_offset))
        self.N = L.orc_num_cells(self.dom)
        self.NB = L.orc_num_bfaces(self.dom)
        self.work = C.c_void_p(L.orc_work_create(self.N))
        self.cfg = StepCfg(corrector_steps, adv_nonortho_steps, p_nonortho_steps, int(nonortho), adv_tol, p_tol, maxit)
        self.bvel = np.ctypeslib.as_array(L.orc_bvel(self.dom), shape=(2, self.NB))

    @classmethod
    def from_compiled(cls, cd, **kw):
        return cls(cd.sizes, cd.btype, cd.bconn, cd.T, cd.bT, cd.bvel0, float(cd.visc),
                   conn_corner_offset=cd.conn_corner_offset, **kw)

    def __del__(self):
        try:
            self.lib.orc_work_free(self.work)
            self.lib.orc_destroy(self.dom)
        except Exception:
            pass

    # work arrays (views)
    def wf(self, which, n):
        return np.ctypeslib.as_array(self.lib.orc_work_field(self.work, which), shape=(n,))

    def build_C(self, u, dt, nonortho=True, for_scalar=False):
        N = self.N
        val = np.zeros((N, 5), np.float32)
        idx = np.zeros((N, 5), np.int32)
        A = np.zeros(N, np.float32)
        u = np.ascontiguousarray(u, np.float32)
        self.lib.orc_build_C(self.dom, _fp(u), C.c_float(dt), int(nonortho), int(for_scalar), _fp(val), _ip(idx), _fp(A))
        return val, idx, A

    def adv_rhs(self, u, ures, dt, nonortho=True):
        rhs = np.zeros((2, self.N), np.float32)
        u = np.ascontiguousarray(u, np.float32)
        ures = np.ascontiguousarray(ures, np.float32)
        self.lib.orc_adv_rhs(self.dom, _fp(u), _fp(ures), C.c_float(dt), int(nonortho), _fp(rhs))
        return rhs

    def build_P(self, A, nonortho=True):
        val = np.zeros((self.N, 5), np.float32)
        idx = np.zeros((self.N, 5), np.int32)
        A = np.ascontiguousarray(A, np.float32)
        self.lib.orc_build_P(self.dom, _fp(A), int(nonortho), _fp(val), _ip(idx))
        return val, idx

    def pressure_rhs(self, u, ures, Cval, Cidx, A, dt):
        h = np.zeros((2, self.N), np.float32)
        args = [np.ascontiguousarray(a, np.float32) for a in (u, ures, Cval)]
        Cidx = np.ascontiguousarray(Cidx, np.int32)
        A = np.ascontiguousarray(A, np.float32)
        self.lib.orc_pressure_rhs(self.dom, _fp(args[0]), _fp(args[1]), _fp(args[2]), _ip(Cidx), _fp(A), C.c_float(dt), _fp(h))
        return h

    def div(self, vel, pres=None, A=None):
        d = np.zeros(self.N, np.float32)
        vel = np.ascontiguousarray(vel, np.float32)
        self.lib.orc_div(self.dom, _fp(vel), _fp(d))
        if pres is not None:
            pres = np.ascontiguousarray(pres, np.float32)
            A = np.ascontiguousarray(A, np.float32)
            self.lib.orc_div_add_nonortho(self.dom, _fp(pres), _fp(A), _fp(d))
        return d

    def correct(self, hbya, p, A):
        out = np.zeros((2, self.N), np.float32)
        a = [np.ascontiguousarray(x, np.float32) for x in (hbya, p, A)]
        self.lib.orc_correct(self.dom, _fp(a[0]), _fp(a[1]), _fp(a[2]), _fp(out))
        return out

    def cg(self, val, idx, f, x0=None, maxit=5000, tol=1e-5, reset=100, best=True):
        x = np.zeros(self.N, np.float32) if x0 is None else np.array(x0, np.float32)
        val = np.ascontiguousarray(val, np.float32)
        idx = np.ascontiguousarray(idx, np.int32)
        f = np.ascontiguousarray(f, np.float32)
        res = C.c_float()
        conv = C.c_int()
        it = self.lib.orc_cg(self.N, _fp(val), _ip(idx), _fp(f), _fp(x), maxit, C.c_float(tol), reset, int(best),
                             C.byref(res), C.byref(conv))
        return x, it, res.value, bool(conv.value)

    def bicgstab(self, val, idx, f, x0=None, maxit=5000, tol=1e-5):
        x = np.zeros(self.N, np.float32) if x0 is None else np.array(x0, np.float32)
        val = np.ascontiguousarray(val, np.float32)
        idx = np.ascontiguousarray(idx, np.int32)
        f = np.ascontiguousarray(f, np.float32)
        res = C.c_float()
        conv = C.c_int()
        it = self.lib.orc_bicgstab(self.N, _fp(val), _ip(idx), _fp(f), _fp(x), maxit, C.c_float(tol), C.byref(res), C.byref(conv))
        return x, it, res.value, bool(conv.value)

    def set_scalar(self, scal_visc, sb_neumann, sbval):
        self.lib.orc_sbval.restype = C.POINTER(C.c_float)
        t = np.ascontiguousarray(sb_neumann, np.int32)
        v = np.ascontiguousarray(sbval, np.float32)
        self.lib.orc_set_scalar(self.dom, C.c_float(scal_visc), _ip(t), _fp(v))
        self.sbval = np.ctypeslib.as_array(self.lib.orc_sbval(self.dom), shape=(self.NB,))

    def scalar_rhs(self, u, T, dt):
        rhs = np.zeros(self.N, np.float32)
        u = np.ascontiguousarray(u, np.float32)
        T = np.ascontiguousarray(T, np.float32)
        self.lib.orc_scalar_rhs(self.dom, _fp(u), _fp(T), C.c_float(dt), _fp(rhs))
        return rhs

    def substep_scalar(self, u, p, T, ures, dt, beta=1.0):
        """Orthogonal-path substep with passive scalar + buoyancy, in place on u, p, T, ures."""
        self.lib.orc_substep_ortho_scalar(self.dom, self.work, C.byref(self.cfg), _fp(u), _fp(p), _fp(T), _fp(ures),
                                          C.c_float(dt), C.c_float(beta))
        return list(self.cfg.bicg_iters), list(self.cfg.cg_iters)[: self.cfg.n_cg], self.cfg.cg_iters[7]

    def sim_step_scalar(self, u, p, T, ures, dt_target, cfl, beta=1.0):
        return self.lib.orc_sim_step_scalar(self.dom, self.work, C.byref(self.cfg), _fp(u), _fp(p), _fp(T), _fp(ures),
                                            C.c_float(dt_target), C.c_float(cfl), C.c_float(beta))

    def max_velocity(self, u):
        u = np.ascontiguousarray(u, np.float32)
        return float(self.lib.orc_max_velocity(self.dom, _fp(u)))

    def flux_balance(self):
        return float(self.lib.orc_flux_balance(self.dom))

    def update_outflow(self, u, out_mask, adj, char_vel, dt, tol):
        u = np.ascontiguousarray(u, np.float32)
        m = np.ascontiguousarray(out_mask, np.uint8)
        adj = np.ascontiguousarray(adj, np.int32)
        cv = np.ascontiguousarray(char_vel, np.float32)
        self.lib.orc_update_outflow(self.dom, _fp(u), m.ctypes.data_as(C.POINTER(C.c_ubyte)), _ip(adj), _fp(cv),
                                    C.c_float(dt), C.c_float(tol))

    def substep(self, u, p, dt):
        """In place on u [2,N], p [N] (float32, contiguous)."""
        assert u.dtype == np.float32 and p.dtype == np.float32 and u.flags.c_contiguous
        self.lib.orc_substep(self.dom, self.work, C.byref(self.cfg), _fp(u), _fp(p), C.c_float(dt))
        return list(self.cfg.bicg_iters), list(self.cfg.cg_iters)[: self.cfg.n_cg]

    def make_divergence_free(self, u, p, maxit=1000):
        self.lib.orc_make_divergence_free(self.dom, self.work, C.byref(self.cfg), _fp(u), _fp(p), maxit)
        return list(self.cfg.cg_iters)[: self.cfg.n_cg]

    def sim_step(self, u, p, dt_target, cfl, out_mask=None, adj=None, char_vel=None, bc_tol=1e-5):
        tcg, tb = C.c_int(0), C.c_int(0)
        if out_mask is not None:
            m = np.ascontiguousarray(out_mask, np.uint8)
            adj = np.ascontiguousarray(adj, np.int32)
            cv = np.ascontiguousarray(char_vel, np.float32)
            n = self.lib.orc_sim_step(self.dom, self.work, C.byref(self.cfg), _fp(u), _fp(p), C.c_float(dt_target),
                                      C.c_float(cfl), m.ctypes.data_as(C.POINTER(C.c_ubyte)), _ip(adj), _fp(cv),
                                      C.c_float(bc_tol), C.byref(tcg), C.byref(tb))
        else:
            n = self.lib.orc_sim_step(self.dom, self.work, C.byref(self.cfg), _fp(u), _fp(p), C.c_float(dt_target),
                                      C.c_float(cfl), None, None, None, C.c_float(bc_tol), C.byref(tcg), C.byref(tb))
        return n, tcg.value, tb.value


def ell_to_csr(val, idx):
    """ELL(5) rows -> CSR sorted by column (the reference's storage, K.cu:3861-3875)."""
    N = val.shape[0]
    mask = idx >= 0
    order = np.argsort(np.where(mask, idx, np.iinfo(np.int32).max), axis=1, kind="stable")
    v = np.take_along_axis(val, order, 1)
    i = np.take_along_axis(idx, order, 1)
    m = np.take_along_axis(mask, order, 1)
    row = np.concatenate([[0], np.cumsum(m.sum(1))]).astype(np.int32)
    return v[m], i[m], row
