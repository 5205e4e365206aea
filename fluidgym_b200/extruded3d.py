"""z-extruded multi-block domains (CylinderJet3D / Airfoil3D: the 2-D multi-block grid repeated over ``nz`` uniform periodic
planes, ``envs/cylinder/grid.py:298``, ``shapes.py:641-676``) -- host side of ``csrc/extruded3_b200.cuh``.

STATUS: the operator arithmetic is verified on the CPU against an op trace of the unmodified reference
(``tests/test_extruded_host.py``, ``tests/test_extruded_cpu.py``); ``ExtrudedPISO3D`` (the launch path) has not run on a GPU yet,
so no environment is registered on it.  ``tools/extruded_check.py`` is the first thing to run on a GPU: one substep from the
reference's traced state, compared with the reference's result.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import native
from .domain import CompiledDomain
from .solver import _TABLE_FIELDS, _ptr


def extruded_neighbours(nbr2: np.ndarray, nz: int) -> np.ndarray:
    """6-face neighbour table [6][nz * N2] of the extruded domain (cell = plane * N2 + g): faces 0..3 are the in-plane
    neighbours of the compiled 2-D domain shifted into the plane (prescribed faces stay negative: no matrix entry), face 4 / 5 the
    same cell in the plane below / above (periodic)."""
    N2 = nbr2.shape[1]
    k = np.arange(nz, dtype=np.int64)[:, None]
    g = np.arange(N2, dtype=np.int64)[None, :]
    out = np.zeros((6, nz, N2), dtype=np.int64)
    for f in range(4):
        n = nbr2[f].astype(np.int64)[None, :]
        out[f] = np.where(n >= 0, n + k * N2, -1)
    out[4] = ((k - 1) % nz) * N2 + g
    out[5] = ((k + 1) % nz) * N2 + g
    return np.ascontiguousarray(out.reshape(6, nz * N2).astype(np.int32))


class ExtrudedPISO3D:
    """State + solver for ``n_envs`` copies of an extruded domain: ``u [B,3,nz*N2]``, ``p [B,nz*N2]``, ``bvel [B,3,nz,NB2]``."""

    def __init__(self, cd: CompiledDomain, nz: int, hz: float, n_envs: int = 1, device="cuda:0", corrector_steps=2,
                 advect_non_ortho_steps=1, pressure_non_ortho_steps=4, advection_tol=1e-5, pressure_tol=5e-7, max_iter=5000):
        if not torch.cuda.is_available():
            raise native.FGBError("fluidgym_b200 needs a CUDA device (there is no CPU fallback)")
        self.lib = native.load()
        self.cd, self.nz, self.hz, self.B = cd, int(nz), float(hz), int(n_envs)
        self.N2, self.NB2, self.N = cd.N, cd.NB, cd.N * int(nz)
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        dev = self.device
        self._tab = {}
        plane = native.Tables()
        plane.N, plane.NB, plane.K_no, plane.K_nob, plane.viscosity = cd.N, cd.NB, cd.K_no, cd.K_nob, float(cd.visc)
        for name in _TABLE_FIELDS:
            arr = getattr(cd, name)
            if name == "b_face":
                arr = arr.astype(np.int8)
            self._tab[name] = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
            setattr(plane, name, self._tab[name].data_ptr())
        self.xtables = native.Extruded3Tables(plane, self.nz, self.hz)
        # Krylov side: an fgb_ortho3 handle on the 6-face table (its metric arrays are not read by the solvers)
        self._tab["nbr6"] = torch.from_numpy(extruded_neighbours(np.asarray(cd.nbr), self.nz)).to(dev)
        self._tab["ones"] = torch.ones(3 * self.N, device=dev)
        t3 = native.Ortho3Tables(self.N, 0, float(cd.visc), self._tab["nbr6"].data_ptr(), self._tab["ones"].data_ptr(),
                                 self._tab["ones"].data_ptr(), self._tab["ones"].data_ptr(), self._tab["ones"].data_ptr(), 0, 0, 0)
        self.tables3 = t3
        self.options = native.Options(corrector_steps, advect_non_ortho_steps, pressure_non_ortho_steps, 1, advection_tol, pressure_tol,
                                      max_iter, 0)
        nbytes = self.lib.fgb_ortho3_workspace_bytes(C.byref(t3), self.B)
        self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        off = (-self.workspace.data_ptr()) % 256
        h = C.c_void_p()
        native.check(self.lib.fgb_ortho3_create(C.byref(t3), self.B, C.c_void_p(self.workspace.data_ptr() + off), nbytes,
                                                C.byref(self.options), C.byref(h)), "fgb_ortho3_create")
        self.handle = h
        self.u = torch.zeros(self.B, 3, self.N, device=dev)
        self.p = torch.zeros(self.B, self.N, device=dev)
        self.bvel = torch.zeros(self.B, 3, self.nz, max(self.NB2, 1), device=dev)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.fgb_ortho3_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def piso_substep(self, dt):
        dtc = dt.to(self.device, torch.float32).contiguous() if isinstance(dt, torch.Tensor) else torch.full((self.B,), float(dt), device=self.device)
        native.check(self.lib.fgb_extruded3_piso_substep(self.handle, C.byref(self.xtables), _ptr(self.u), _ptr(self.p), _ptr(self.bvel),
                                                         _ptr(dtc), self.stream), "fgb_extruded3_piso_substep")
