"""TEST INFRASTRUCTURE ONLY.  CPU stand-in for ``fluidgym_b200.extruded3d.ExtrudedPISO3D``: the same state layout and the same
``ExtrudedStepping`` host logic, with the substep executed by the ``__host__ __device__`` per-cell code of
``fluidgym_b200/csrc/extruded3_b200.cuh`` compiled for the host (tests/cpu_harness/extruded_host.cu) and numpy Krylov solvers in
place of the cooperative CUDA ones.  The call sequence mirrors ``fgb_extruded3_piso_substep`` /
``fgb_extruded3_make_divergence_free`` statement by statement.  Never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp
import torch

from fluidgym_b200 import native
from fluidgym_b200.extruded3d import ExtrudedStepping, extruded_neighbours

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
f32 = np.float32
FIELDS = ["nbr", "fl_comp", "minv", "det", "Cd", "Wp", "no_idx", "no_face", "no_gP", "no_gN", "no_wv", "nob_idx", "nob_w", "b_minv", "b_det",
          "b_alpha", "b_cell", "b_face"]


def build_harness():
    src = os.path.join(ROOT, "tests", "cpu_harness", "extruded_host.cu")
    out = os.path.join(ROOT, "tests", "_build", "libextruded_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    hdr = os.path.join(ROOT, "fluidgym_b200", "csrc", "extruded3_b200.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.check_call([nvcc, "-x", "cu", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared",
                               "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "fluidgym_b200", "csrc"), "-o", out, src])
    return C.CDLL(out)


def host_tables(cd):
    """fgb_tables with HOST pointers into the numpy arrays of a compiled 2-D domain"""
    keep = {}
    t = native.Tables()
    t.N, t.NB, t.K_no, t.K_nob, t.viscosity = cd.N, cd.NB, cd.K_no, cd.K_nob, float(cd.visc)
    for name in FIELDS:
        a = np.ascontiguousarray(getattr(cd, name))
        if name == "b_face":
            a = np.ascontiguousarray(a.astype(np.int8))
        keep[name] = a
        setattr(t, name, a.ctypes.data)
    return t, keep


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def csr(nbr6, coff, diag):
    n = diag.size
    rows, cols, vals = [np.arange(n)], [np.arange(n)], [diag.reshape(-1)]
    for f in range(6):
        ok = nbr6[f] >= 0
        rows.append(np.nonzero(ok)[0]); cols.append(nbr6[f][ok]); vals.append(coff[f].reshape(-1)[ok])
    return sp.csr_matrix((np.concatenate(vals).astype(f32), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


def bicgstab(M, b, tol, maxit=500):
    """unpreconditioned BiCGStab, zero start, stop on ||r|| / sqrt(N) < tol (BICG.cu:237-376)"""
    n = b.size
    x = np.zeros(n, f32); r = b.copy(); rw = r.copy(); p = r.copy()
    rho = alpha = omega = f32(1)
    v = np.zeros(n, f32)
    for i in range(maxit):
        if np.sqrt(np.dot(r, r)) / np.sqrt(n) < tol:
            return x, i
        rho_new = np.dot(rw, r)
        if i > 0:
            p = r + (rho_new / rho) * (alpha / omega) * (p - omega * v)
        rho = rho_new
        v = M @ p
        alpha = rho / np.dot(rw, v)
        x = x + alpha * p
        r = r - alpha * v
        if np.sqrt(np.dot(r, r)) / np.sqrt(n) < tol:
            return x, i + 1
        tt = M @ r
        omega = np.dot(tt, r) / np.dot(tt, tt)
        x = x + omega * r
        r = r - omega * tt
    return x, maxit


def cg(M, b, x0, tol, reset=100, maxit=5000):
    """CG with residual reset, best-iterate return and mean removal (CG.cu:225-446, SIM.py:1908-1925)"""
    n = b.size
    x = x0.copy()
    if not b.any():
        return np.zeros(n, f32), 0
    r = b - M @ x
    p = r.copy()
    rho = np.dot(r, r)
    it = 0
    best, best_rr = x, np.inf
    for i in range(maxit):
        if reset and (i + 1) % reset == 0:
            r = b - M @ x; p = r.copy(); rho = np.dot(r, r)
        Ap = M @ p
        alpha = rho / np.dot(p, Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        rr = np.dot(r, r)
        it = i + 1
        if rr < best_rr:
            best, best_rr = x, rr
        if np.sqrt(rr) / np.sqrt(n) < tol:
            break
        p = r + (rr / rho) * p
        rho = rr
    x = best
    return (x - x.mean()).astype(f32), it


class HostExtrudedPISO3D(ExtrudedStepping):
    def __init__(self, cd, nz, hz, n_envs=1, device="cpu", corrector_steps=2, advect_non_ortho_steps=1, pressure_non_ortho_steps=4,
                 advection_tol=1e-5, pressure_tol=5e-7, max_iter=5000):
        self.lib = build_harness()
        self.cd, self.nz, self.hz, self.B = cd, int(nz), float(hz), int(n_envs)
        self.N2, self.NB2, self.N = cd.N, cd.NB, cd.N * int(nz)
        self.device = torch.device("cpu")
        self.t, self._keep = host_tables(cd)
        self.nbr6 = extruded_neighbours(np.asarray(cd.nbr), self.nz)
        self.opt = dict(cs=corrector_steps, ans=advect_non_ortho_steps, pns=pressure_non_ortho_steps, atol=advection_tol, ptol=pressure_tol,
                        max_iter=max_iter)
        self.u = torch.zeros(self.B, 3, self.N)
        self.p = torch.zeros(self.B, self.N)
        self.bvel = torch.zeros(self.B, 3, self.nz, self.NB2)
        self.cg_iters, self.bicg_iters = [], []

    def _shape(self):
        return self.nz, self.N2

    def piso_substep(self, dt):
        dtv = dt.numpy() if isinstance(dt, torch.Tensor) else np.full(self.B, float(dt), f32)
        for b in range(self.B):
            self._substep_env(b, float(dtv[b]))

    def _substep_env(self, b, dt):
        L, t, o = self.lib, self.t, self.opt
        nz, N2 = self._shape()
        hzc, dtc = C.c_float(self.hz), C.c_float(dt)
        u = np.ascontiguousarray(self.u[b].numpy().reshape(3, nz, N2))
        p = np.ascontiguousarray(self.p[b].numpy().reshape(nz, N2)).copy()
        bvel = np.ascontiguousarray(self.bvel[b].numpy())
        coff, A, rhs = np.zeros((6, nz, N2), f32), np.zeros((nz, N2), f32), np.zeros((3, nz, N2), f32)
        ures = np.zeros((3, nz, N2), f32)
        Cm = None
        for ns in range(o["ans"]):
            L.xh_setup_advection(C.byref(t), nz, hzc, _p(u), _p(u if ns == 0 else ures), _p(bvel), dtc, _p(coff), _p(A), _p(rhs), int(ns == 0))
            if Cm is None:
                Cm = csr(self.nbr6, coff, A)
            nxt = np.zeros_like(ures)
            for c in range(3):
                x, it = bicgstab(Cm, rhs[c].reshape(-1), o["atol"])
                nxt[c] = x.reshape(nz, N2)
                self.bicg_iters.append(it)
            ures = nxt
        poff, pdiag = np.zeros((6, nz, N2), f32), np.zeros((nz, N2), f32)
        hb, div = np.zeros((3, nz, N2), f32), np.zeros((nz, N2), f32)
        Pm = None
        for cs in range(o["cs"]):
            if cs == 0:
                L.xh_pressure_matrix(C.byref(t), nz, hzc, _p(A), _p(poff), _p(pdiag))
                Pm = csr(self.nbr6, poff, pdiag)
            L.xh_hbya(C.byref(t), nz, hzc, _p(u), _p(ures), _p(bvel), _p(coff), _p(A), dtc, _p(hb))
            for ps in range(o["pns"]):
                L.xh_divergence(C.byref(t), nz, hzc, _p(hb), _p(bvel), _p(p), _p(A), _p(div))
                x, it = cg(Pm, div.reshape(-1), np.zeros(nz * N2, f32) if ps == 0 else p.reshape(-1), o["ptol"], 100, o["max_iter"])
                p = np.ascontiguousarray(x.reshape(nz, N2))
                self.cg_iters.append(it)
            out = np.zeros((3, nz, N2), f32)
            L.xh_correct(C.byref(t), nz, hzc, _p(hb), _p(p), _p(A), _p(out))
            ures = out
        self.u[b] = torch.from_numpy(ures.reshape(3, -1))
        self.p[b] = torch.from_numpy(p.reshape(-1))

    def make_divergence_free(self, max_iter=1000):
        L, t, o = self.lib, self.t, self.opt
        nz, N2 = self._shape()
        hzc = C.c_float(self.hz)
        for b in range(self.B):
            u = np.ascontiguousarray(self.u[b].numpy().reshape(3, nz, N2))
            p = np.ascontiguousarray(self.p[b].numpy().reshape(nz, N2)).copy()
            bvel = np.ascontiguousarray(self.bvel[b].numpy())
            A = np.ones((nz, N2), f32)
            poff, pdiag, div = np.zeros((6, nz, N2), f32), np.zeros((nz, N2), f32), np.zeros((nz, N2), f32)
            L.xh_pressure_matrix(C.byref(t), nz, hzc, _p(A), _p(poff), _p(pdiag))
            Pm = csr(self.nbr6, poff, pdiag)
            for ps in range(o["pns"]):
                L.xh_divergence(C.byref(t), nz, hzc, _p(u), _p(bvel), _p(p), _p(A), _p(div))
                x, it = cg(Pm, div.reshape(-1), np.zeros(nz * N2, f32) if ps == 0 else p.reshape(-1), o["ptol"], 0, max_iter)
                p = np.ascontiguousarray(x.reshape(nz, N2))
            out = np.zeros((3, nz, N2), f32)
            L.xh_correct(C.byref(t), nz, hzc, _p(u), _p(p), _p(A), _p(out))
            self.u[b] = torch.from_numpy(out.reshape(3, -1))
            self.p[b] = torch.from_numpy(p.reshape(-1))
