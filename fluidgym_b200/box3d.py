"""3-D single-block rectilinear ("box") domains and their batched PISO solver (turbulent channel flow).

Host-side mirror of what the reference does with ``PISOtorch.Domain(3, ...)`` + one ``CreateBlock`` whose axes are
periodic or closed (``envs/tcf/grid.py:166-272``), compiled into the flat tables of ``fgb_ortho3_tables``
(``include/fluidgym_b200.h``).  Metrics follow ``grid_gen.cu:298-354`` (cell: face-centre differences) and
``:398-494`` (boundary faces: one-sided), restricted to their diagonal -- exact on rectilinear grids.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import native
from .solver import _ptr

f32 = np.float32


def _face_centres(v, axis, side):
    """centres of the faces normal to ``axis`` (0 = x) on the lower / upper side of every cell; v: [3, nz+1, ny+1, nx+1]"""
    ax = 3 - axis
    sl = [slice(None)] * 4
    sl[ax] = slice(1, None) if side else slice(0, -1)
    a = v[tuple(sl)]
    o0, o1 = [i for i in (1, 2, 3) if i != ax]
    s = 0
    for i0 in (0, 1):
        for i1 in (0, 1):
            s2 = [slice(None)] * 4
            s2[o0] = slice(1, None) if i0 else slice(0, -1)
            s2[o1] = slice(1, None) if i1 else slice(0, -1)
            s = s + a[tuple(s2)]
    return (s * f32(0.25)).astype(f32)


class Box3DDomain:
    """vertex [3, nz+1, ny+1, nx+1] float32 (x, y, z coordinates); closed[d]: axis d (0 = x) has Dirichlet walls,
    otherwise it is periodic.  Cell g = x + nx (y + ny z); prescribed faces are numbered face-major (-x, +x, -y, ...)
    with the tangential cells in (z, y, x) order -- the layout of ``FixedBoundary.velocity`` flattened."""

    def __init__(self, vertex, closed=(False, True, False), viscosity=1e-3, transforms=None, btransforms=None):
        """transforms [nz,ny,nx,19] / btransforms {face: [...,19]}: optional externally computed metric tensors
        (M row-major, M^-1 row-major, det -- the layout of Block.transform) used instead of the own ones; the parity
        tests pass the reference's so that operator comparisons start from bit-identical inputs."""
        v = np.ascontiguousarray(vertex, dtype=f32)
        self.vertex = v
        self.nz, self.ny, self.nx = (s - 1 for s in v.shape[1:])
        self.shape = (self.nz, self.ny, self.nx)
        self.N = self.nx * self.ny * self.nz
        self.closed = tuple(bool(c) for c in closed)
        self.visc = float(f32(viscosity))
        h = np.stack([(_face_centres(v, d, 1)[d] - _face_centres(v, d, 0)[d]).astype(f32) for d in range(3)])
        det = (h[0] * h[1] * h[2]).astype(f32)
        r = (f32(1.0) / det).astype(f32)
        self.h, self.det = h, det
        self.minv = np.stack([h[1] * h[2] * r, h[0] * h[2] * r, h[0] * h[1] * r]).astype(f32)
        if transforms is not None:
            T = np.asarray(transforms, dtype=f32).reshape(self.shape + (19,))
            self.h = np.stack([T[..., 0], T[..., 4], T[..., 8]])
            self.det = det = np.ascontiguousarray(T[..., 18])
            self.minv = np.stack([T[..., 9], T[..., 13], T[..., 17]]).astype(f32)
        cells = np.arange(self.N, dtype=np.int64).reshape(self.shape)
        nbr = np.zeros((6,) + self.shape, dtype=np.int64)
        self.boff, self.bshape = {}, {}
        b_minv, b_det, nb = [], [], 0
        for f in range(6):
            d, up = f >> 1, f & 1
            ax = 2 - d
            nbr[f] = np.roll(cells, -1 if up else 1, axis=ax)
            if self.closed[d]:
                sl = [slice(None)] * 3
                sl[ax] = -1 if up else 0
                layer = tuple(sl)
                n_face = cells[layer].size
                nbr[f][layer] = -1 - (nb + np.arange(n_face).reshape(cells[layer].shape))
                self.boff[f], self.bshape[f] = nb, cells[layer].shape
                # boundary-face transform = the adjacent cell's (one-sided difference over the same half cell pair)
                if btransforms is not None and f in btransforms:
                    bt = np.asarray(btransforms[f], dtype=f32).reshape(-1, 19)
                    b_minv.append(np.stack([bt[:, 9], bt[:, 13], bt[:, 17]]))
                    b_det.append(bt[:, 18])
                else:
                    b_minv.append(np.stack([self.minv[k][layer].ravel() for k in range(3)]))
                    b_det.append(det[layer].ravel())
                nb += n_face
        self.NB = nb
        self.nbr = np.ascontiguousarray(nbr.reshape(6, self.N).astype(np.int32))
        self.b_minv = np.ascontiguousarray(np.concatenate(b_minv, axis=1)) if nb else np.zeros((3, 1), f32)
        self.b_det = np.ascontiguousarray(np.concatenate(b_det)) if nb else np.zeros(1, f32)

    def cell_centres(self):
        v = self.vertex
        c = 0
        for i in (0, 1):
            for j in (0, 1):
                for k in (0, 1):
                    c = c + v[:, i:v.shape[1] - 1 + i, j:v.shape[2] - 1 + j, k:v.shape[3] - 1 + k]
        return (c * f32(0.125)).astype(f32)


class BatchedPISO3D:
    """State + solver for ``n_envs`` copies of one Box3DDomain: ``u [B,3,N]``, ``p [B,N]``, ``bvel [B,3,NB]``."""

    def __init__(self, dom: Box3DDomain, n_envs: int = 1, device="cuda:0", corrector_steps=2, advection_tol=1e-6, pressure_tol=1e-6,
                 max_iter=5000):
        if not torch.cuda.is_available():
            raise native.FGBError("fluidgym_b200 needs a CUDA device (there is no CPU fallback)")
        self.lib = native.load()
        self.dom, self.B, self.N, self.NB = dom, int(n_envs), dom.N, dom.NB
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        dev = self.device
        self._tab = {k: torch.from_numpy(np.ascontiguousarray(getattr(dom, k))).to(dev) for k in ("nbr", "b_minv", "b_det")}
        self._tab["minv"] = torch.from_numpy(np.ascontiguousarray(dom.minv.reshape(3, -1))).to(dev)
        self._tab["det"] = torch.from_numpy(np.ascontiguousarray(dom.det.reshape(-1))).to(dev)
        self.tables = native.Ortho3Tables(dom.N, dom.NB, dom.visc, *[self._tab[k].data_ptr() for k in ("nbr", "minv", "det", "b_minv", "b_det")])
        self.options = native.Options(corrector_steps, 1, 1, 1, advection_tol, pressure_tol, max_iter, 0)
        nbytes = self.lib.fgb_ortho3_workspace_bytes(C.byref(self.tables), self.B)
        self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        off = (-self.workspace.data_ptr()) % 256
        h = C.c_void_p()
        native.check(self.lib.fgb_ortho3_create(C.byref(self.tables), self.B, C.c_void_p(self.workspace.data_ptr() + off), nbytes,
                                                C.byref(self.options), C.byref(h)), "fgb_ortho3_create")
        self.handle = h
        B, N, NB = self.B, self.N, max(self.NB, 1)
        self.u = torch.zeros(B, 3, N, device=dev)
        self.p = torch.zeros(B, N, device=dev)
        self.bvel = torch.zeros(B, 3, NB, device=dev)

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

    def buffer(self, name):
        B, N = self.B, self.N
        shapes = {"Coff": ((B, 6, N), torch.float32), "A": ((B, N), torch.float32), "rhs": ((B, 3, N), torch.float32),
                  "ures": ((B, 3, N), torch.float32), "Poff": ((B, 6, N), torch.float32), "Pdiag": ((B, N), torch.float32),
                  "hbya": ((B, 3, N), torch.float32), "div": ((B, N), torch.float32), "iters": ((B, 8), torch.int32),
                  "resid": ((B, 8), torch.float32), "dt": ((B,), torch.float32), "nsub": ((B,), torch.int32), "maxvel": ((B,), torch.float32),
                  "src": ((B, 4), torch.float32), "rowmean": ((B, 4), torch.float32), "iter_total": ((B, 2), torch.int64)}
        shape, dtype = shapes[name]
        ptr = self.lib.fgb_ortho3_buffer(self.handle, name.encode())
        o = ptr - self.workspace.data_ptr()
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self.workspace[o:o + n].view(dtype).view(shape)

    def _dt(self, dt):
        if isinstance(dt, torch.Tensor):
            return dt.to(self.device, torch.float32).contiguous()
        return torch.full((self.B,), float(dt), device=self.device)

    # ---- per-op (names follow the PISOtorch free functions) ------------------------------------------------------
    def setup_advection(self, dt, src=None):
        self._dtc = self._dt(dt)
        native.check(self.lib.fgb_ortho3_setup_advection(self.handle, _ptr(self.u), _ptr(self.bvel), _ptr(src), _ptr(self._dtc), None,
                                                         self.stream), "fgb_ortho3_setup_advection")

    def solve_advection(self, zero_init=True):
        native.check(self.lib.fgb_ortho3_solve_advection(self.handle, int(zero_init), None, self.stream), "fgb_ortho3_solve_advection")

    def setup_pressure(self, dt, src=None, with_matrix=True):
        self._dtc = self._dt(dt)
        native.check(self.lib.fgb_ortho3_setup_pressure(self.handle, _ptr(self.u), _ptr(self.bvel), _ptr(src), _ptr(self._dtc),
                                                        int(with_matrix), None, self.stream), "fgb_ortho3_setup_pressure")

    def solve_pressure(self, p_out=None, zero_init=True, reset_steps=100, max_iter=None, slot=0):
        p_out = self.p if p_out is None else p_out
        native.check(self.lib.fgb_ortho3_solve_pressure(self.handle, _ptr(p_out), int(zero_init), reset_steps,
                                                        max_iter or self.options.max_iter, slot, None, self.stream), "fgb_ortho3_solve_pressure")

    def correct_velocity(self, p=None, u_out=None):
        p = self.p if p is None else p
        u_out = self.buffer("ures") if u_out is None else u_out
        native.check(self.lib.fgb_ortho3_correct_velocity(self.handle, _ptr(p), _ptr(u_out), None, self.stream), "fgb_ortho3_correct_velocity")

    # ---- fused ---------------------------------------------------------------------------------------------------
    def piso_substep(self, dt, src=None):
        self._dtc = self._dt(dt)
        native.check(self.lib.fgb_ortho3_piso_substep(self.handle, _ptr(self.u), _ptr(self.p), _ptr(self.bvel), _ptr(src), _ptr(self._dtc),
                                                      None, self.stream), "fgb_ortho3_piso_substep")

    def make_divergence_free(self, max_iter=1000):
        native.check(self.lib.fgb_ortho3_make_divergence_free(self.handle, _ptr(self.u), _ptr(self.p), _ptr(self.bvel), max_iter,
                                                              self.stream), "fgb_ortho3_make_divergence_free")

    def single_step(self, dt, cfl, rows=None, d_lo=1.0, d_hi=1.0) -> int:
        n = C.c_int32(0)
        n_row = 0 if rows is None else int(rows.shape[1])
        native.check(self.lib.fgb_ortho3_sim_step(self.handle, _ptr(self.u), _ptr(self.p), _ptr(self.bvel), float(dt), float(cfl),
                                                  _ptr(rows), n_row, float(d_lo), float(d_hi), C.byref(n), self.stream), "fgb_ortho3_sim_step")
        return n.value

    def wall_rows(self, rows, d_lo, d_hi, set_forcing=False, acc=None):
        native.check(self.lib.fgb_ortho3_wall_rows(self.handle, _ptr(self.u), _ptr(rows), int(rows.shape[1]), float(d_lo), float(d_hi),
                                                   int(set_forcing), _ptr(acc), self.stream), "fgb_ortho3_wall_rows")
        return self.buffer("rowmean")
