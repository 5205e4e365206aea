"""TEST INFRASTRUCTURE ONLY (part of the CPU oracle; never imported by the product package fluidgym_b200).

numpy evaluation of the compiled domain tables (test helper, CPU only).

States, in vectorised numpy, exactly what each CUDA kernel in fluidgym_b200/csrc computes from the
tables of ``fluidgym_b200.domain.CompiledDomain``.  Used by the CPU test-suite to check the table
compiler against the literal oracle (oracle/piso_oracle.c) without a GPU.
"""
import numpy as np

f32 = np.float32


def contravariant(cd, u):
    """U^k = det * (Minv[k] . u) per cell -> [2, N]"""
    mi, det = cd.minv, cd.det
    return np.stack([det * (mi[0] * u[0] + mi[1] * u[1]), det * (mi[2] * u[0] + mi[3] * u[1])]).astype(f32)


def boundary_flux(cd, bvel):
    """F_b = det_b * (Minv_b[axis] . u_b) per boundary face -> [NB]"""
    ax = cd.b_face >> 1
    r0 = np.where(ax == 0, cd.b_minv[0], cd.b_minv[2])
    r1 = np.where(ax == 0, cd.b_minv[1], cd.b_minv[3])
    return (cd.b_det * (r0 * bvel[0] + r1 * bvel[1])).astype(f32)


def face_fluxes(cd, vel, bvel):
    """[4, N] face fluxes (not multiplied by the face sign), K.cu:1567-1645"""
    U = contravariant(cd, vel)
    Fb = boundary_flux(cd, bvel)
    out = np.zeros((4, cd.N), f32)
    for f in range(4):
        nb = cd.nbr[f]
        inner = nb >= 0
        comp = (cd.fl_comp[f] & 1).astype(np.int64)
        sgn = np.where(cd.fl_comp[f] & 2, f32(-1), f32(1))
        j = np.where(inner, nb, 0)
        velN = sgn * U[comp, j]
        out[f] = np.where(inner, (velN + U[f >> 1]) * f32(0.5), Fb[np.where(inner, 0, -1 - nb)])
    return out


def assemble_C(cd, u, bvel, dt):
    """ELL off-diagonals [4, N] and diagonal A [N] of the advection-diffusion matrix (rows / det)."""
    fl = face_fluxes(cd, u, bvel)
    diag = cd.det / f32(dt) + cd.Cd[0]
    off = np.zeros((4, cd.N), f32)
    for f in range(4):
        inner = cd.nbr[f] >= 0
        ff = f32((f & 1) * 2 - 1) * f32(0.5) * fl[f]
        diag = diag + np.where(inner, ff, 0)
        off[f] = np.where(inner, (ff + cd.Cd[f + 1]) / cd.det, 0)
    return off.astype(f32), (diag / cd.det).astype(f32)


def boundary_source(cd, bvel):
    """[2, N] Dirichlet boundary advection + diffusion sources (before / det), K.cu:4321-4380"""
    Fb = boundary_flux(cd, bvel)
    S = np.zeros((2, cd.N), f32)
    for f in range(4):
        nb = cd.nbr[f]
        bnd = nb < 0
        j = np.where(bnd, -1 - nb, 0)
        fs = f32((f & 1) * 2 - 1)
        for c in range(2):
            vel = bvel[c, j]
            S[c] += np.where(bnd, -vel * (Fb[j] * fs) + vel * cd.visc * 2 * cd.b_alpha[j], 0)
    return S


def nonortho_velocity(cd, field, bvel_c):
    S = np.zeros(cd.N, f32)
    for k in range(cd.K_no):
        S += cd.no_wv[k] * field[cd.no_idx[k]]
    for k in range(cd.K_nob):
        S += cd.nob_w[k] * bvel_c[cd.nob_idx[k]]
    return S


def adv_rhs(cd, u, ures, bvel, dt, src=None):
    bs = boundary_source(cd, bvel)
    rhs = np.zeros((2, cd.N), f32)
    for c in range(2):
        r = cd.det * u[c] / f32(dt) + bs[c] - nonortho_velocity(cd, ures[c], bvel[c])
        r = r / cd.det
        if src is not None:
            r = r + src[c]
        rhs[c] = r
    return rhs


def neighbor_values(cd, x):
    """[5, N]: own value and the 4 face neighbours (own value where the face is prescribed)."""
    out = [x]
    for f in range(4):
        nb = cd.nbr[f]
        out.append(np.where(nb >= 0, x[np.where(nb >= 0, nb, 0)], x))
    return np.stack(out)


def build_P(cd, A):
    rA = neighbor_values(cd, (f32(1) / A).astype(f32))
    P = np.einsum("ejn,jn->en", cd.Wp, rA).astype(f32)
    return P[1:], P[0]


def hbya(cd, u, ures, Coff, A, bvel, dt, src=None):
    bs = boundary_source(cd, bvel)
    out = np.zeros((2, cd.N), f32)
    for c in range(2):
        H = np.zeros(cd.N, f32)
        for f in range(4):
            nb = cd.nbr[f]
            H += np.where(nb >= 0, Coff[f] * ures[c][np.where(nb >= 0, nb, 0)], 0)
        S = bs[c] / cd.det
        if src is not None:
            S = S + src[c]
        out[c] = (u[c] / f32(dt) - H + S) / A
    return out


def divergence(cd, vel, bvel, pres=None, A=None):
    fl = face_fluxes(cd, vel, bvel)
    d = (fl[1] - fl[0]) + (fl[3] - fl[2])
    if pres is not None:
        rA = neighbor_values(cd, (f32(1) / A).astype(f32))
        S = np.zeros(cd.N, f32)
        for k in range(cd.K_no):
            w = cd.no_gP[k] * rA[0] + cd.no_gN[k] * rA[1 + cd.no_face[k].astype(np.int64), np.arange(cd.N)]
            S += w * pres[cd.no_idx[k]]
        d = d + S
    return d.astype(f32)


def correct(cd, hb, p, A):
    pv = neighbor_values(cd, p)
    fac = [np.where((cd.nbr[2 * d] < 0) | (cd.nbr[2 * d + 1] < 0), f32(1), f32(0.5)) for d in range(2)]
    g0 = (pv[2] - pv[1]) * fac[0]
    g1 = (pv[4] - pv[3]) * fac[1]
    gx = g0 * cd.minv[0] + g1 * cd.minv[2]
    gy = g0 * cd.minv[1] + g1 * cd.minv[3]
    rA = f32(1) / A
    return np.stack([hb[0] - rA * gx, hb[1] - rA * gy]).astype(f32)


def spmv(cd, off, diag, x):
    y = diag * x
    for f in range(4):
        nb = cd.nbr[f]
        y = y + np.where(nb >= 0, off[f] * x[np.where(nb >= 0, nb, 0)], 0)
    return y.astype(f32)


def max_velocity(cd, u, bvel):
    mi = cd.minv
    a = np.abs(mi[0] * u[0] + mi[1] * u[1]).max()
    b = np.abs(mi[2] * u[0] + mi[3] * u[1]).max()
    bm = cd.b_minv
    c = np.abs(bm[0] * bvel[0] + bm[1] * bvel[1]).max()
    d = np.abs(bm[2] * bvel[0] + bm[3] * bvel[1]).max()
    return float(max(a, b, c, d))


def scalar_setup(cd, u, T, bvel, sbval, dt):
    """Scalar transport matrix (off [4,N], diag [N]) and right-hand side, as k_setup_scalar computes them."""
    fl = face_fluxes(cd, u, bvel)
    diag = cd.det / f32(dt) + cd.Cd_s[0]
    r = cd.det * T / f32(dt)
    off = np.zeros((4, cd.N), f32)
    kap = cd.scalar_visc
    for f in range(4):
        nb = cd.nbr[f]
        inner = nb >= 0
        fs = f32((f & 1) * 2 - 1)
        ff = fs * f32(0.5) * fl[f]
        diag = diag + np.where(inner, ff, 0)
        off[f] = np.where(inner, (ff + cd.Cd_s[f + 1]) / cd.det, 0)
        j = np.where(inner, 0, -1 - nb)
        sc = sbval[j]
        visc = np.where(cd.sb_neumann[j] == 0, sc * kap * 2 * cd.b_alpha[j], sc * kap)
        r = r + np.where(inner, 0, -sc * (fl[f] * fs) + visc)
    return off.astype(f32), (diag / cd.det).astype(f32), (r / cd.det).astype(f32)
