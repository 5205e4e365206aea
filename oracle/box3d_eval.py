"""TEST INFRASTRUCTURE ONLY (part of the CPU oracle; never imported by the product package fluidgym_b200).

float32 numpy statement of the PISO substep on a 3-D single-block ORTHOGONAL box (periodic or wall-bounded per
axis) -- the specification of the D = 3 kernels (fluidgym_b200/csrc/piso3d_b200.cu), written against the same
reference formulas as oracle/table_eval.py (K.cu:495-537 contravariant fluxes, :1224-1466 Laplace coefficients,
:3617-3880 C, :4296-4400 RHS, :4812-4978 P, :5136-5255 HbyA, :5389-5434 divergence, :816-849 + :5962-5995
corrector) with the metric tensors restricted to their diagonal: on a rectilinear grid every off-diagonal
(non-orthogonal) coefficient is exactly zero, so the deferred corrections vanish (`if (alpha != 0)`, K.cu:3772).

Layout: cell g = x + nx*(y + ny*z); fields [3, N]; faces 0..5 = -x,+x,-y,+y,-z,+z; boundary faces of a closed axis
are numbered face-major, tangential index (slow, fast) = the two remaining axes in (z,y,x) order.
"""
import numpy as np

f32 = np.float32


class Box3D:
    def __init__(self, vertex, closed=(False, True, False), viscosity=1e-3, T=None, bT=None):
        """vertex [3, nz+1, ny+1, nx+1] float32; closed[d]: axis d has fixed (Dirichlet) boundaries, else periodic."""
        v = np.asarray(vertex, dtype=f32)
        self.nz, self.ny, self.nx = (s - 1 for s in v.shape[1:])
        self.shape = (self.nz, self.ny, self.nx)
        self.N = self.nx * self.ny * self.nz
        self.closed = tuple(bool(c) for c in closed)
        self.visc = f32(viscosity)
        h, det = cell_metrics(v) if T is None else (np.stack([T[..., 0], T[..., 4], T[..., 8]]), T[..., 18])
        self.h = h.astype(f32)                              # [3, nz, ny, nx] diagonal of M
        self.det = det.astype(f32)
        if T is None:
            r = (f32(1.0) / self.det).astype(f32)
            self.minv = np.stack([h[1] * h[2] * r, h[0] * h[2] * r, h[0] * h[1] * r]).astype(f32)
        else:
            self.minv = np.stack([T[..., 9], T[..., 13], T[..., 17]]).astype(f32)
        self.alpha = (self.det * self.minv * self.minv).astype(f32)          # [3, ...] det * |M^-1 row d|^2
        # boundary faces: per closed axis d, lower and upper: metrics [2 sides][...] on the face layer
        self.b = {}
        for d in range(3):
            if not self.closed[d]:
                continue
            for side in (0, 1):
                f = 2 * d + side
                if bT is not None:
                    t = bT[f]
                    bh = np.stack([t[..., 0], t[..., 4], t[..., 8]])
                    bdet = t[..., 18]
                    bminv = np.stack([t[..., 9], t[..., 13], t[..., 17]])
                else:
                    bh, bdet = boundary_metrics(v, d, side)
                    r = (f32(1.0) / bdet).astype(f32)
                    bminv = np.stack([bh[1] * bh[2] * r, bh[0] * bh[2] * r, bh[0] * bh[1] * r]).astype(f32)
                self.b[f] = dict(det=bdet.astype(f32), minv=bminv.astype(f32), alpha=(bdet * bminv[d] * bminv[d]).astype(f32))

    # -- helpers ------------------------------------------------------------------------------------
    def field(self, a):
        return np.asarray(a, dtype=f32).reshape((-1,) + self.shape)

    def shift(self, a, d, s):
        """value of the neighbour in direction s (+1/-1) along axis d (array axes are z,y,x -> axis 2-d)."""
        return np.roll(a, -s, axis=a.ndim - 1 - d)

    def inner(self, f):
        """mask [nz,ny,nx]: face f of the cell has a neighbour cell (not a fixed boundary)."""
        d, up = f >> 1, f & 1
        m = np.ones(self.shape, dtype=bool)
        if self.closed[d]:
            idx = [slice(None)] * 3
            idx[2 - d] = -1 if up else 0
            m[tuple(idx)] = False
        return m

    def bexpand(self, f, vals):
        """boundary-face array [..., on the face layer] -> full cell array that is non-zero only in the boundary layer"""
        d, up = f >> 1, f & 1
        out = np.zeros(vals.shape[:-3] + self.shape, dtype=f32)
        idx = [slice(None)] * out.ndim
        idx[out.ndim - 1 - d] = slice(-1, None) if up else slice(0, 1)
        out[tuple(idx)] = vals
        return out


def cell_metrics(v):
    """diagonal of M (face-centre differences, grid_gen.cu:298-354) and det, [3, nz, ny, nx]"""
    q = f32(0.25)

    def fc(axis, side):          # centre of the faces normal to `axis` (array axis 3-axis... v is [3, z, y, x])
        a = v
        sl = [slice(None)] * 4
        ax = 3 - axis
        sl[ax] = slice(1, None) if side else slice(0, -1)
        a = a[tuple(sl)]
        others = [i for i in (1, 2, 3) if i != ax]
        s = 0
        for i0 in (0, 1):
            for i1 in (0, 1):
                sl2 = [slice(None)] * 4
                sl2[others[0]] = slice(1, None) if i0 else slice(0, -1)
                sl2[others[1]] = slice(1, None) if i1 else slice(0, -1)
                s = s + a[tuple(sl2)]
        return (s * q).astype(f32)
    h = np.stack([(fc(d, 1)[d] - fc(d, 0)[d]).astype(f32) for d in range(3)])
    det = (h[0] * h[1] * h[2]).astype(f32)
    return h, det


def boundary_metrics(v, d, side):
    """face transform on the boundary layer of axis d (one-sided normal distance, grid_gen.cu:398-494)"""
    h, det = cell_metrics(v)
    sl = [slice(None)] * 4
    sl[3 - d] = slice(-1, None) if side else slice(0, 1)
    bh = h[tuple(sl)].copy()
    return bh, (bh[0] * bh[1] * bh[2]).astype(f32)


# ---- operators -------------------------------------------------------------------------------------------
def face_fluxes(g: Box3D, u, bvel):
    """[6, nz,ny,nx] contravariant face fluxes; bvel: {face: [3, face layer]} Dirichlet velocities"""
    u = g.field(u)
    U = (g.det * g.minv * u).astype(f32)          # U^d = det * minv_d * u_d
    fl = np.zeros((6,) + g.shape, dtype=f32)
    for f in range(6):
        d, up = f >> 1, f & 1
        fl[f] = (f32(0.5) * (U[d] + g.shift(U[d], d, 1 if up else -1))).astype(f32)
        if g.closed[d]:
            b = g.b[f]
            Fb = (b["det"] * b["minv"][d] * bvel[f][d]).astype(f32)
            m = ~g.inner(f)
            fl[f] = np.where(m, g.bexpand(f, Fb), fl[f])
    return fl


def assemble(g: Box3D, u, bvel, dt, visc=None):
    """off [6, ...], A [...] of C/det (K.cu:3617-3880).  visc: optional per-cell viscosity [nz, ny, nx] (block viscosity set by a prep
    function, e.g. the Smagorinsky model): face coefficient (alpha_P nu_P + alpha_N nu_N) / 2, wall term 2 alpha_P nu_P (K.cu:3697-3750, 3845)"""
    fl = face_fluxes(g, u, bvel)
    diag = (g.det / f32(dt)).astype(f32)
    off = np.zeros((6,) + g.shape, dtype=f32)
    nuP = g.visc if visc is None else np.asarray(visc, f32).reshape(g.shape)
    for f in range(6):
        d, up = f >> 1, f & 1
        sig = f32(1.0 if up else -1.0)
        inner = g.inner(f)
        aN = g.shift(g.alpha[d], d, 1 if up else -1)
        nuN = g.visc if visc is None else g.shift(nuP, d, 1 if up else -1)
        vc = ((g.alpha[d] * nuP + aN * nuN) * f32(0.5)).astype(f32)
        ff = (sig * f32(0.5) * fl[f]).astype(f32)
        diag = diag + np.where(inner, ff + vc, f32(2.0) * nuP * g.alpha[d])
        off[f] = np.where(inner, (ff - vc) / g.det, 0).astype(f32)
    return off.astype(f32), (diag / g.det).astype(f32), fl


def boundary_source(g: Box3D, bvel, fl, visc=None):
    """visc: optional per-cell viscosity; the wall term uses the viscosity of the adjacent cell (getViscosityFixedBoundary, K.cu:1840-1843)"""
    Sb = np.zeros((3,) + g.shape, dtype=f32)
    for f in range(6):
        d, up = f >> 1, f & 1
        if not g.closed[d]:
            continue
        sig = f32(1.0 if up else -1.0)
        b = g.b[f]
        Fb = (b["det"] * b["minv"][d] * bvel[f][d]).astype(f32)
        if visc is None:
            nu_b = g.visc
        else:
            sl = [slice(None)] * 3
            sl[2 - d] = -1 if up else 0
            nu_b = np.asarray(visc, f32).reshape(g.shape)[tuple(sl)].reshape(b["alpha"].shape)
        k = (-(sig * Fb) + f32(2.0) * nu_b * b["alpha"]).astype(f32)
        Sb += g.bexpand(f, (bvel[f] * k).astype(f32))
    return Sb


def adv_rhs(g: Box3D, u, bvel, dt, src=None, visc=None):
    u = g.field(u)
    fl = face_fluxes(g, u, bvel)
    Sb = boundary_source(g, bvel, fl, visc)
    rhs = ((g.det * u / f32(dt) + Sb) / g.det).astype(f32)
    if src is not None:
        src = np.asarray(src, dtype=f32)
        rhs = rhs + (src.reshape(3, 1, 1, 1) if src.size == 3 else src.reshape((3,) + g.shape))
    return rhs.astype(f32), Sb


def spmv(g: Box3D, off, diag, x):
    x = g.field(x)
    y = diag * x
    for f in range(6):
        y = y + off[f] * g.shift(x, f >> 1, 1 if (f & 1) else -1)
    return y.astype(f32)


def build_P(g: Box3D, A):
    rA = (f32(1.0) / A).astype(f32)
    off = np.zeros((6,) + g.shape, dtype=f32)
    diag = np.zeros(g.shape, dtype=f32)
    for f in range(6):
        d, up = f >> 1, f & 1
        s = 1 if up else -1
        c = (f32(0.5) * (g.alpha[d] * rA + g.shift(g.alpha[d], d, s) * g.shift(rA, d, s))).astype(f32)
        c = np.where(g.inner(f), c, 0).astype(f32)
        off[f] = c
        diag = diag - c
    return off, diag.astype(f32)


def hbya(g: Box3D, u, ures, off, A, Sb, dt, src=None):
    u, ures = g.field(u), g.field(ures)
    H = np.zeros_like(u)
    for f in range(6):
        H = H + off[f] * g.shift(ures, f >> 1, 1 if (f & 1) else -1)
    inner = u / f32(dt) - H + Sb / g.det
    if src is not None:
        src = np.asarray(src, dtype=f32)
        inner = inner + (src.reshape(3, 1, 1, 1) if src.size == 3 else src.reshape((3,) + g.shape))
    return (inner / A).astype(f32)


def divergence(g: Box3D, vel, bvel):
    fl = face_fluxes(g, vel, bvel)
    return ((fl[1] - fl[0]) + (fl[3] - fl[2]) + (fl[5] - fl[4])).astype(f32)


def correct(g: Box3D, hb, p, A):
    p = g.field(p)[0]
    out = np.zeros((3,) + g.shape, dtype=f32)
    for d in range(3):
        pu = np.where(g.inner(2 * d + 1), g.shift(p, d, 1), p)
        pl = np.where(g.inner(2 * d), g.shift(p, d, -1), p)
        fac = np.where(g.inner(2 * d + 1) & g.inner(2 * d), f32(0.5), f32(1.0))
        out[d] = hb[d] - (pu - pl) * fac * g.minv[d] / A
    return out.astype(f32)


# ---- passive scalar + buoyancy (RBC3D) ---------------------------------------------------------------------
def assemble_scalar(g: Box3D, u, bvel, T, sbval, dt, kappa):
    """off [6, ...], A, rhs of the scalar transport (K.cu:3617-3880 with forPassiveScalar, K.cu:4094-4198);
    sbval: {face: [face layer]} Dirichlet values on the closed faces."""
    kappa = f32(kappa)
    fl = face_fluxes(g, u, bvel)
    T = g.field(T)[0]
    diag = (g.det / f32(dt)).astype(f32)
    r = (g.det * T / f32(dt)).astype(f32)
    off = np.zeros((6,) + g.shape, dtype=f32)
    for f in range(6):
        d, up = f >> 1, f & 1
        sig = f32(1.0 if up else -1.0)
        inner = g.inner(f)
        aN = g.shift(g.alpha[d], d, 1 if up else -1)
        vc = ((g.alpha[d] * kappa + aN * kappa) * f32(0.5)).astype(f32)
        ff = (sig * f32(0.5) * fl[f]).astype(f32)
        diag = diag + np.where(inner, ff + vc, f32(2.0) * kappa * g.alpha[d])
        off[f] = np.where(inner, (ff - vc) / g.det, 0).astype(f32)
        if g.closed[d]:
            b = g.b[f]
            Fb = (b["det"] * b["minv"][d] * bvel[f][d]).astype(f32)
            r = r + g.bexpand(f, (sbval[f] * (-(sig * Fb) + f32(2.0) * kappa * b["alpha"])).astype(f32))     # sbval[f]: face layer
    return off.astype(f32), (diag / g.det).astype(f32), (r / g.det).astype(f32)


def buoyancy_source(g: Box3D, T, beta):
    """velocity source field (0, beta T, 0) [3, nz, ny, nx] (rbc_env_base.py:280-304)"""
    T = g.field(T)[0]
    z = np.zeros_like(T)
    return np.stack([z, (f32(beta) * T).astype(f32), z])
