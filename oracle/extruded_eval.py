"""TEST INFRASTRUCTURE ONLY (part of the CPU oracle; never imported by the product package fluidgym_b200).

float32 numpy statement of the PISO operators on a z-EXTRUDED multi-block domain (CylinderJet3D, Airfoil3D: the 2-D
multi-block grid repeated over nz uniform, periodic z planes; envs/cylinder/grid.py:298, shapes.py:641-676) -- test helper,
CPU only, the specification for the D = 3 non-orthogonal kernels that are the next row of SURVEY section 8(f).

The metric tensor of an extruded cell is block diagonal, M3 = diag(M2, hz): det3 = hz det2, alpha3^{ij} = hz alpha2^{ij} in
the plane, alpha3^{zz} = det2 / hz and alpha3^{xz} = alpha3^{yz} = 0, so every non-orthogonal (corner) term of the reference's
DIMS = 3 kernels lives in the x-y plane (`if (alpha != 0)`, K.cu:3772).  Rows are divided by det3, hence all in-plane
coefficients equal the 2-D ones of oracle/table_eval.py (compiled tables of fluidgym_b200.domain.CompiledDomain) and the
z faces add  +-1/4 (u_z,P + u_z,N) / hz - nu / hz^2  off the diagonal and  2 nu / hz^2 + 1/2 (F_z+ - F_z-) / (det2 hz)  on it.
Validated against an op trace of the unmodified reference on CylinderJet3D-easy (resolution 8, 5 blocks x 8 planes,
tests/golden/cyl3d_substep0.npz) in tests/test_extruded_cpu.py.

Layout: fields [C, nz, N2] with N2 the block-major 2-D cell index of the compiled domain; boundary values [C, nz, NB2].
"""
import numpy as np

import table_eval as te

f32 = np.float32


def _zn(a, s):
    """value in the z-neighbour plane k + s (periodic); planes are axis -2"""
    return np.roll(a, -s, axis=-2)


def assemble(cd, u, bvel, dt, hz):
    """in-plane off-diagonals [4, nz, N2], z off-diagonals [2, nz, N2] (-z, +z) and A [nz, N2] of C / det3 (K.cu:3617-3880)"""
    nz = u.shape[1]
    hz = f32(hz)
    off = np.zeros((4, nz, cd.N), f32)
    A = np.zeros((nz, cd.N), f32)
    for k in range(nz):
        off[:, k], A[k] = te.assemble_C(cd, u[:2, k], bvel[:2, k], dt)
    uz = u[2]
    visc = f32(cd.visc)
    Fzp = f32(0.5) * (uz + _zn(uz, 1))            # F_z / det2
    Fzm = f32(0.5) * (uz + _zn(uz, -1))
    dz = visc / (hz * hz)
    offz = np.stack([(-f32(0.5) * Fzm) / hz - dz, (f32(0.5) * Fzp) / hz - dz]).astype(f32)
    A = (A + f32(2) * dz + f32(0.5) * (Fzp - Fzm) / hz).astype(f32)
    return off, offz, A


def spmv(cd, off, offz, diag, x):
    y = np.zeros_like(x)
    for k in range(x.shape[0]):
        y[k] = te.spmv(cd, off[:, k], diag[k], x[k])
    return (y + offz[0] * _zn(x, -1) + offz[1] * _zn(x, 1)).astype(f32)


def adv_rhs(cd, u, ures, bvel, dt):
    """predictor right-hand side for the three components (K.cu:4296-4400): the 2-D boundary sources and deferred
    non-orthogonal terms act on every component within its plane"""
    nz = u.shape[1]
    rhs = np.zeros_like(u)
    for k in range(nz):
        bs = te.boundary_source(cd, bvel[:2, k])            # uses the in-plane boundary flux; per-component values follow
        Fb = te.boundary_flux(cd, bvel[:2, k])
        for c in range(3):
            if c < 2:
                S = bs[c]
            else:                                           # z component: same formula with the z boundary values
                S = np.zeros(cd.N, f32)
                for f in range(4):
                    nb = cd.nbr[f]
                    bnd = nb < 0
                    j = np.where(bnd, -1 - nb, 0)
                    fs = f32((f & 1) * 2 - 1)
                    vel = bvel[2, k][j]
                    S += np.where(bnd, -vel * (Fb[j] * fs) + vel * cd.visc * 2 * cd.b_alpha[j], 0)
            r = cd.det * u[c, k] / f32(dt) + S - te.nonortho_velocity(cd, ures[c, k], bvel[c, k])
            rhs[c, k] = r / cd.det
    return rhs.astype(f32)


def build_P(cd, A, hz):
    """pressure matrix (K.cu:4812-4978), NOT divided by det: in-plane weights * hz, z faces 1/2 (det2 / hz) (1/A_P + 1/A_N)"""
    hz = f32(hz)
    nz = A.shape[0]
    off = np.zeros((4, nz, cd.N), f32)
    diag = np.zeros((nz, cd.N), f32)
    for k in range(nz):
        o, d = te.build_P(cd, A[k])
        off[:, k], diag[k] = o * hz, d * hz
    rA = (f32(1) / A).astype(f32)
    az = cd.det / hz
    offz = np.stack([f32(0.5) * az * (rA + _zn(rA, -1)), f32(0.5) * az * (rA + _zn(rA, 1))]).astype(f32)
    return off, offz, (diag - offz[0] - offz[1]).astype(f32)


def hbya(cd, u, ures, off, offz, A, bvel, dt):
    nz = u.shape[1]
    out = np.zeros_like(u)
    for k in range(nz):
        bs = te.boundary_source(cd, bvel[:2, k])
        Fb = te.boundary_flux(cd, bvel[:2, k])
        for c in range(3):
            H = np.zeros(cd.N, f32)
            for f in range(4):
                nb = cd.nbr[f]
                H += np.where(nb >= 0, off[f, k] * ures[c, k][np.where(nb >= 0, nb, 0)], 0)
            H = H + offz[0, k] * ures[c, (k - 1) % nz] + offz[1, k] * ures[c, (k + 1) % nz]
            if c < 2:
                S = bs[c]
            else:
                S = np.zeros(cd.N, f32)
                for f in range(4):
                    nb = cd.nbr[f]
                    bnd = nb < 0
                    j = np.where(bnd, -1 - nb, 0)
                    fs = f32((f & 1) * 2 - 1)
                    vel = bvel[2, k][j]
                    S += np.where(bnd, -vel * (Fb[j] * fs) + vel * cd.visc * 2 * cd.b_alpha[j], 0)
            out[c, k] = (u[c, k] / f32(dt) - H + S / cd.det) / A[k]
    return out.astype(f32)


def divergence(cd, vel, bvel, hz, pres=None, A=None):
    """flux divergence (K.cu:5389-5434) + the deferred non-orthogonal pressure term (K.cu:5470-5492); fluxes are the 3-D
    ones: in-plane 2-D fluxes * hz, z faces det2 * mean(u_z)"""
    hz = f32(hz)
    nz = vel.shape[1]
    d = np.zeros((nz, cd.N), f32)
    for k in range(nz):
        d[k] = te.divergence(cd, vel[:2, k], bvel[:2, k], None if pres is None else pres[k], None if A is None else A[k]) * hz
    uz = vel[2]
    return (d + cd.det * (f32(0.5) * (uz + _zn(uz, 1)) - f32(0.5) * (uz + _zn(uz, -1)))).astype(f32)


def correct(cd, hb, p, A, hz):
    hz = f32(hz)
    nz = p.shape[0]
    out = np.zeros_like(hb)
    for k in range(nz):
        out[:2, k] = te.correct(cd, hb[:2, k], p[k], A[k])
    out[2] = hb[2] - (f32(1) / A) * (f32(0.5) * (_zn(p, 1) - _zn(p, -1)) / hz)
    return out.astype(f32)


# ---- boundary hooks of the 3-D cylinder (jets + outflow) --------------------------------------------------------------
def _flux_weights(cd):
    """signed in-plane boundary-flux weights [2, NB2] per unit plane spacing: F_b = hz * (fw[0] u_b + fw[1] v_b)"""
    NB = cd.NB
    face = cd.b_face[:NB].astype(np.int64)
    ax = face >> 1
    sign = np.where(face & 1, 1.0, -1.0).astype(f32)
    j = np.arange(NB)
    bm, bd = cd.b_minv[:, :NB], cd.b_det[:NB]
    return np.stack([bd * bm[2 * ax, j] * sign, bd * bm[2 * ax + 1, j] * sign]).astype(f32)


def balance_fluxes(cd, bvel, free, hz, tol):
    """balance_boundary_fluxes (SIM.py:188-224) on [3, nz, NB2] boundary values: the velocities of the ``free`` faces (bool [NB2],
    all planes) are scaled by -(flux through the other prescribed faces) / (flux through the free faces) unless the imbalance is
    below 0.01 tol."""
    fw = _flux_weights(cd)
    fl = (bvel[0] * fw[0] + bvel[1] * fw[1]) * f32(hz)
    fixed, var = fl[:, ~free].sum(dtype=np.float64), fl[:, free].sum(dtype=np.float64)
    out = bvel.copy()
    if not abs(fixed + var) <= tol * 0.01:
        out[:, :, free] *= f32(-fixed / var)
    return out


def apply_jets(cd, bvel, control, jet_faces, jet_templ, out_mask, hz):
    """CylinderJetEnv3D._apply_action (jet_cylinder_env_3d.py:399-424): control [nz] (the smoothed action of the jet that owns
    the plane), jet faces take template * control in every plane (z component 0), then the jets AND the outflow are rescaled for a
    zero net boundary flux (tol 1e-7)."""
    out = bvel.copy()
    for k in range(out.shape[1]):
        out[:2, k][:, jet_faces] = jet_templ * f32(control[k])
        out[2, k][jet_faces] = 0
    free = out_mask.copy()
    free[jet_faces] = True
    return balance_fluxes(cd, out, free, hz, 1e-7)


def update_outflow(cd, u, bvel, dt, out_mask, hz, char_vel=(1.0, 0.0), tol=1e-5):
    """update_advective_boundaries + balance_boundary_fluxes (SIM.py:188-393) of the "PRE" hook: relaxation of the outflow values
    towards the adjacent cell with weight 1 - 1 / (1 + 2 dt U_adv) for all three components, then the outflow alone is rescaled."""
    NB = cd.NB
    o = np.nonzero(out_mask)[0]
    ax = (cd.b_face[:NB].astype(np.int64) >> 1)[o]
    bm = cd.b_minv[:, :NB]
    adv = bm[2 * ax, o] * f32(char_vel[0]) + bm[2 * ax + 1, o] * f32(char_vel[1])
    w = f32(1) - f32(1) / (f32(1) + f32(2) * f32(dt) * adv)
    cells = cd.b_cell[:NB][o]
    out = bvel.copy()
    for c in range(3):
        out[c][:, o] = out[c][:, o] - w * (out[c][:, o] - u[c][:, cells])
    return balance_fluxes(cd, out, out_mask, hz, tol)
