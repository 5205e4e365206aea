"""3-D single-block rectilinear ("box") domains and their batched PISO solver (turbulent channel flow).

Host-side mirror of what the reference does with ``PISOtorch.Domain(3, ...)`` + one ``CreateBlock`` whose axes are
periodic or closed (``envs/tcf/grid.py:166-272``), compiled into the flat tables of ``fgb_ortho3_tables``
(``include/fluidgym_b200.h``).  Metrics follow ``grid_gen.cu:298-354`` (cell: face-centre differences) and
``:398-494`` (boundary faces: one-sided), restricted to their diagonal -- exact on rectilinear grids.
"""
from __future__ import annotations

import ctypes as C
import os

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


def structured_neighbours(nx, ny, nz, closed, halo=False, boff=None):
    """The neighbour arithmetic of the kernels (``o3_nbrs``, csrc/ortho3_b200.cuh) on the (z, y, x) cell ordering: [6, N]
    indices, -1 - j for the prescribed face j (``boff``: first face index per cell face; all -1 without it), halo planes at
    N.. / N + P.. for slabs.  Host mirror used by the tests to pin the formula to the tables (``Box3DDomain.nbr`` / ``SlabTables.nbr``)."""
    P, N = nx * ny, nx * ny * nz
    g = np.arange(N, dtype=np.int64)
    q = g // nx
    i, j, k = g % nx, q % ny, g // P
    cx, cy, cz = (bool(c) for c in closed)

    def face(f, idx):
        return -1 - (boff[f] + idx) if boff is not None else np.full(N, -1, dtype=np.int64)
    n = np.empty((6, N), dtype=np.int64)
    n[0] = np.where(i > 0, g - 1, face(0, q) if cx else g + (nx - 1))
    n[1] = np.where(i < nx - 1, g + 1, face(1, q) if cx else g - (nx - 1))
    n[2] = np.where(j > 0, g - nx, face(2, k * nx + i) if cy else g + (ny - 1) * nx)
    n[3] = np.where(j < ny - 1, g + nx, face(3, k * nx + i) if cy else g - (ny - 1) * nx)
    n[4] = np.where(k > 0, g - P, N + (g - k * P) if halo else (face(4, g - k * P) if cz else g + (nz - 1) * P))
    n[5] = np.where(k < nz - 1, g + P, N + P + (g - k * P) if halo else (face(5, g - k * P) if cz else g - (nz - 1) * P))
    return n


class BatchedPISO3D:
    """State + solver for ``n_envs`` copies of one Box3DDomain: ``u [B,3,N]``, ``p [B,N]``, ``bvel [B,3,NB]``."""

    def __init__(self, dom: Box3DDomain, n_envs: int = 1, device="cuda:0", corrector_steps=2, advection_tol=1e-6, pressure_tol=1e-6,
                 max_iter=5000, non_orthogonal=True):
        """non_orthogonal: the reference's code path (Simulation(non_orthogonal=...)): True = zero-started predictor and CG with
        residual reset every 100 iterations (TCF), False = predictor started from the previous result, no reset (RBC)."""
        if not torch.cuda.is_available():
            raise native.FGBError("fluidgym_b200 needs a CUDA device (there is no CPU fallback)")
        self.lib = native.load_for(device)
        self.dom, self.B, self.N, self.NB = dom, int(n_envs), dom.N, dom.NB
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        dev = self.device
        self._tab = {k: torch.from_numpy(np.ascontiguousarray(getattr(dom, k))).to(dev) for k in ("nbr", "b_minv", "b_det")}
        self._tab["minv"] = torch.from_numpy(np.ascontiguousarray(dom.minv.reshape(3, -1))).to(dev)
        self._tab["det"] = torch.from_numpy(np.ascontiguousarray(dom.det.reshape(-1))).to(dev)
        self.tables = native.Ortho3Tables(dom.N, dom.NB, dom.visc, *[self._tab[k].data_ptr() for k in ("nbr", "minv", "det", "b_minv", "b_det")])
        if os.environ.get("FGB_O3_BOX", "1") != "0":          # structured neighbour arithmetic in the Krylov kernels (0: nbr table)
            self.tables.nx, self.tables.ny, self.tables.nz = dom.nx, dom.ny, dom.nz
            self.tables.closed = sum(1 << d for d in range(3) if dom.closed[d])
            for f, o in dom.boff.items():
                self.tables.boff[f] = int(o)
        self.options = native.Options(corrector_steps, 1, 1, int(bool(non_orthogonal)), advection_tol, pressure_tol, max_iter, 0)
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
        self.has_scalar = False

    # ---- passive scalar + buoyancy (RBC3D) ---------------------------------------------------------------------------
    def attach_scalar(self, kappa: float, beta: float = 1.0, sbval0=None):
        """Domain(passiveScalarChannels=1) + setScalarViscosity (rbc_env_base.py:206-216): adds ``T [B,N]`` and the Dirichlet
        boundary values ``sbval [B,NB]``; every substep then transports T with the incoming velocity first and adds the
        buoyancy source (0, beta T, 0)."""
        dev = self.device
        self.T = torch.zeros(self.B, self.N, device=dev)
        self.sbval = torch.zeros(self.B, max(self.NB, 1), device=dev)
        if sbval0 is not None:
            self.sbval.copy_(torch.as_tensor(np.asarray(sbval0, dtype=np.float32), device=dev).expand_as(self.sbval))
        self.kappa, self.beta = float(kappa), float(beta)
        self.has_scalar = True
        self._scalar = native.Ortho3Scalar(self.T.data_ptr(), self.sbval.data_ptr(), self.kappa, self.beta)
        native.check(self.lib.fgb_ortho3_set_scalar(self.handle, C.byref(self._scalar)), "fgb_ortho3_set_scalar")

    # ---- sub-grid-scale viscosity + velocity gradients ---------------------------------------------------------------------
    def set_sgs(self, coefficient: float, damping=None):
        """Smagorinsky model (tcf_env.py:441-472): per-cell viscosity nu + C delta |S| damping, refreshed before every substep.
        ``damping``: [N] squared van Driest factor (envs/tcf/grid.py:101-125) or None; coefficient 0 switches the model off."""
        self._sgs_damp = None if damping is None else torch.as_tensor(np.asarray(damping, dtype=np.float32).reshape(-1), device=self.device).contiguous()
        if self._sgs_damp is not None and self._sgs_damp.numel() != self.N:
            raise ValueError(f"set_sgs: damping has {self._sgs_damp.numel()} entries for {self.N} cells")
        native.check(self.lib.fgb_ortho3_set_sgs(self.handle, float(coefficient), _ptr(self._sgs_damp)), "fgb_ortho3_set_sgs")

    def sgs_viscosity(self):
        """PISOtorch.SGSviscosityIncompressibleSmagorinsky + the prep function's damping and base viscosity -> ``visc [B,N]``"""
        native.check(self.lib.fgb_ortho3_sgs_viscosity(self.handle, _ptr(self.u), _ptr(self.bvel), None, self.stream), "fgb_ortho3_sgs_viscosity")
        return self.buffer("visc")

    def velocity_gradients(self):
        """PISOtorch.ComputeSpatialVelocityGradients: ``[B, 3 (component c), 3 (direction d), N]`` = d u_c / d x_d; ``out[:, i]`` is the
        i-th tensor of the reference's list (K.cu:6470-6480: list index = component, channel = direction)"""
        out = torch.empty(self.B, 3, 3, self.N, device=self.device)
        native.check(self.lib.fgb_ortho3_velocity_gradients(self.handle, _ptr(self.u), _ptr(self.bvel), _ptr(out), self.stream),
                     "fgb_ortho3_velocity_gradients")
        return out

    def q_criterion(self):
        """Q = (|Omega|^2 - |S|^2) / 2 (tcf_env.py:586-644, before the reference's resampling to the output grid) -> [B, N]"""
        g = self.velocity_gradients()                                # [B, c, d, N]
        s, o = 0.5 * (g + g.transpose(1, 2)), 0.5 * (g - g.transpose(1, 2))
        return 0.5 * ((o * o).sum(dim=(1, 2)) - (s * s).sum(dim=(1, 2)))

    def advect_scalar(self, dt):
        self._dtc = self._dt(dt)
        native.check(self.lib.fgb_ortho3_advect_scalar(self.handle, _ptr(self.u), _ptr(self.bvel), _ptr(self._dtc), None, self.stream),
                     "fgb_ortho3_advect_scalar")

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
                  "hbya": ((B, 3, N), torch.float32), "div": ((B, N), torch.float32), "visc": ((B, N), torch.float32), "iters": ((B, 8), torch.int32),
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


# ---------------------------------------------------------------------------------------------------------------------
# slab decomposition of one box over the GPUs of a node (BASELINE config 5): one process per GPU, z-slabs
# ---------------------------------------------------------------------------------------------------------------------
class _RawCuda:
    """exposes a raw device pointer to torch through ``__cuda_array_interface__``"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class SlabTables:
    """Tables of z-slab ``rank`` of ``world`` of a Box3DDomain: owned cells [0, N) in (z_local, y, x) order, lower halo plane
    [N, N + P), upper halo plane [N + P, N + 2P) (P = nx ny); the -z / +z neighbours of the first / last owned plane are the
    halo cells, whose metrics are copied from the neighbouring slabs (the grid is static)."""

    def __init__(self, dom: Box3DDomain, rank: int, world: int):
        if dom.closed[2]:
            raise ValueError("slab decomposition runs along z, which must be periodic")
        if dom.nz % world or dom.nz // world < 2:
            raise ValueError(f"nz = {dom.nz} must be a multiple of the world size {world} with at least 2 planes per rank")
        self.dom, self.rank, self.world = dom, rank, world
        nx, ny, nz = dom.nx, dom.ny, dom.nz
        self.nzl = nz // world
        self.z0 = rank * self.nzl
        P = self.P = nx * ny
        N = self.N = P * self.nzl
        NS = self.NS = N + 2 * P
        zs = slice(self.z0, self.z0 + self.nzl)
        zl, zu = (self.z0 - 1) % nz, (self.z0 + self.nzl) % nz
        cells = np.arange(N, dtype=np.int64).reshape(self.nzl, ny, nx)
        nbr = np.zeros((6, self.nzl, ny, nx), dtype=np.int64)
        self.boff, nb = {}, 0
        b_minv, b_det = [], []
        for f in range(6):
            d, up = f >> 1, f & 1
            ax = 2 - d
            nbr[f] = np.roll(cells, -1 if up else 1, axis=ax)
            if d == 2:
                off = np.arange(P).reshape(ny, nx)
                if up:
                    nbr[f][-1] = N + P + off
                else:
                    nbr[f][0] = N + off
            elif dom.closed[d]:
                sl = [slice(None)] * 3
                sl[ax] = -1 if up else 0
                layer = tuple(sl)
                n_face = cells[layer].size
                nbr[f][layer] = -1 - (nb + np.arange(n_face).reshape(cells[layer].shape))
                self.boff[f] = nb
                gl = dom.boff[f]
                gshape = dom.bshape[f]                       # global face layer shape (z, tangential)
                gm = dom.b_minv[:, gl:gl + int(np.prod(gshape))].reshape((3,) + gshape)
                gd = dom.b_det[gl:gl + int(np.prod(gshape))].reshape(gshape)
                b_minv.append(gm[:, zs].reshape(3, -1))
                b_det.append(gd[zs].reshape(-1))
                nb += n_face
        self.NB = nb
        self.nbr = np.full((6, NS), 0, dtype=np.int32)
        self.nbr[:, :N] = nbr.reshape(6, N)

        def with_halo(a):                                   # [nz, ny, nx] global -> [NS]
            return np.concatenate([a[zs].reshape(-1), a[zl].reshape(-1), a[zu].reshape(-1)]).astype(f32)
        self.minv = np.stack([with_halo(dom.minv[k]) for k in range(3)])
        self.det = with_halo(dom.det)
        self.b_minv = np.ascontiguousarray(np.concatenate(b_minv, axis=1)) if nb else np.zeros((3, 1), f32)
        self.b_det = np.ascontiguousarray(np.concatenate(b_det)) if nb else np.zeros(1, f32)

    def take_cells(self, a):
        """[..., N_global] -> [..., NS] (owned part filled, halos zero)"""
        a = np.asarray(a)
        g = a.reshape(a.shape[:-1] + (self.dom.nz, self.P))[..., self.z0:self.z0 + self.nzl, :].reshape(a.shape[:-1] + (self.N,))
        out = np.zeros(a.shape[:-1] + (self.NS,), dtype=a.dtype)
        out[..., :self.N] = g
        return out

    def take_faces(self, a, f):
        """boundary values of global face f [..., nz * tangential] -> this slab's faces"""
        gshape = self.dom.bshape[f]
        a = np.asarray(a)
        lead = a.shape[:-1]
        a = a.reshape(lead + gshape)
        return a[..., self.z0:self.z0 + self.nzl, :].reshape(lead + (-1,))


class SlabPISO3D:
    """One rank of the slab-decomposed solver.  ``torch.distributed`` is used once, at construction, to exchange the CUDA IPC
    handles of the symmetric regions; the solver path itself contains no collective call."""

    PAD_BYTES = 4096

    def __init__(self, dom: Box3DDomain, rank: int, world: int, device, group=None, corrector_steps=2, advection_tol=1e-6,
                 pressure_tol=1e-6, max_iter=5000):
        import torch.distributed as dist
        self.lib = native.load_for(device)
        self.tabs = tb = SlabTables(dom, rank, world)
        self.rank, self.world = rank, world
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        dev = self.device
        self._tab = {k: torch.from_numpy(np.ascontiguousarray(getattr(tb, k))).to(dev) for k in ("nbr", "minv", "det", "b_minv", "b_det")}
        self.tables = native.Ortho3Tables(tb.N, tb.NB, dom.visc, *[self._tab[k].data_ptr() for k in ("nbr", "minv", "det", "b_minv", "b_det")],
                                          tb.NS, dom.N, tb.P)
        if os.environ.get("FGB_O3_BOX", "1") != "0":
            self.tables.nx, self.tables.ny, self.tables.nz = dom.nx, dom.ny, tb.nzl
            self.tables.closed = sum(1 << d for d in range(2) if dom.closed[d])      # z: slab halo planes
            for f, o in tb.boff.items():
                self.tables.boff[f] = int(o)
        self.options = native.Options(corrector_steps, 1, 1, 1, advection_tol, pressure_tol, max_iter, 0)
        ws_bytes = self.lib.fgb_ortho3_workspace_bytes(C.byref(self.tables), 1)

        def al(n):
            return (n + 255) // 256 * 256
        NB = max(tb.NB, 1)
        self._off = {"u": self.PAD_BYTES}
        self._off["p"] = self._off["u"] + al(3 * tb.NS * 4)
        self._off["bvel"] = self._off["p"] + al(tb.NS * 4)
        self._off["ws"] = self._off["bvel"] + al(3 * NB * 4)
        total = self._off["ws"] + al(ws_bytes)
        base, handle = C.c_void_p(), C.create_string_buffer(64)
        native.check(self.lib.fgb_ipc_alloc(total, C.byref(base), handle), "fgb_ipc_alloc")
        self.base, self.total = base.value, total
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        peers = (C.c_void_p * world)()
        self._opened = []
        for q in range(world):
            if q == rank:
                peers[q] = self.base
            else:
                pp = C.c_void_p()
                native.check(self.lib.fgb_ipc_open(handles[q], C.byref(pp)), "fgb_ipc_open")
                peers[q] = pp.value
                self._opened.append(pp.value)
        raw = torch.as_tensor(_RawCuda(self.base, total), device=dev)
        self._raw = raw

        def view(name, shape):
            n = int(np.prod(shape)) * 4
            return raw[self._off[name]:self._off[name] + n].view(torch.float32).view(shape)
        self.u, self.p, self.bvel = view("u", (1, 3, tb.NS)), view("p", (1, tb.NS)), view("bvel", (1, 3, NB))
        h = C.c_void_p()
        native.check(self.lib.fgb_ortho3_create(C.byref(self.tables), 1, C.c_void_p(self.base + self._off["ws"]), al(ws_bytes),
                                                C.byref(self.options), C.byref(h)), "fgb_ortho3_create")
        self.handle = h
        native.check(self.lib.fgb_ortho3_set_slab(h, rank, world, C.c_void_p(self.base), peers), "fgb_ortho3_set_slab")
        dist.barrier(group=group)          # every rank has mapped every region before the first kernel writes to a peer

    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def buffer(self, name):
        shapes = {"iters": ((1, 8), torch.int32), "resid": ((1, 8), torch.float32), "rowmean": ((1, 4), torch.float32),
                  "src": ((1, 4), torch.float32), "dt": ((1,), torch.float32), "maxvel": ((1,), torch.float32),
                  "iter_total": ((1, 2), torch.int64), "ures": ((1, 3, self.tabs.NS), torch.float32), "A": ((1, self.tabs.NS), torch.float32)}
        shape, dtype = shapes[name]
        ptr = self.lib.fgb_ortho3_buffer(self.handle, name.encode())
        o = ptr - self.base
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self._raw[o:o + n].view(dtype).view(shape)

    def load_global(self, u, p, bvel_faces=None):
        """u [3, N_global], p [N_global] numpy; bvel_faces {face: [3, n_face_global]}"""
        tb = self.tabs
        self.u.copy_(torch.from_numpy(tb.take_cells(np.asarray(u, dtype=f32))).to(self.device).unsqueeze(0))
        self.p.copy_(torch.from_numpy(tb.take_cells(np.asarray(p, dtype=f32))).to(self.device).unsqueeze(0))
        self.bvel.zero_()
        for f, v in (bvel_faces or {}).items():
            loc = tb.take_faces(np.asarray(v, dtype=f32), f)
            self.bvel[0, :, tb.boff[f]:tb.boff[f] + loc.shape[-1]] = torch.from_numpy(np.ascontiguousarray(loc)).to(self.device)

    def owned(self, t):
        return t[..., :self.tabs.N]

    def set_sgs(self, coefficient: float, damping_global=None):
        """as BatchedPISO3D.set_sgs; ``damping_global``: [N_global] (this rank keeps its slab's cells)"""
        self._sgs_damp = None
        if damping_global is not None:
            loc = self.tabs.take_cells(np.asarray(damping_global, dtype=f32).reshape(-1))[..., :self.tabs.N]
            self._sgs_damp = torch.from_numpy(np.ascontiguousarray(loc)).to(self.device)
        native.check(self.lib.fgb_ortho3_set_sgs(self.handle, float(coefficient), _ptr(self._sgs_damp)), "fgb_ortho3_set_sgs")

    def piso_substep(self, dt, src=None):
        dtc = torch.full((1,), float(dt), device=self.device)
        native.check(self.lib.fgb_ortho3_piso_substep(self.handle, _ptr(self.u), _ptr(self.p), _ptr(self.bvel), _ptr(src), _ptr(dtc), None,
                                                      self.stream), "fgb_ortho3_piso_substep")

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

    def error(self) -> int:
        e = C.c_int32(0)
        native.check(self.lib.fgb_ortho3_slab_error(self.handle, C.byref(e)), "fgb_ortho3_slab_error")
        return e.value

    def close(self):
        if getattr(self, "handle", None):
            torch.cuda.synchronize(self.device)
            self.lib.fgb_ortho3_destroy(self.handle)
            self.handle = None
            for pp in self._opened:
                self.lib.fgb_ipc_close(C.c_void_p(pp))
            del self.u, self.p, self.bvel, self._raw
            self.lib.fgb_ipc_free(C.c_void_p(self.base))
