"""TEST INFRASTRUCTURE ONLY (part of the CPU oracle; never imported by the product package fluidgym_b200).

float64 numpy statement of one PISO substep on the compiled tables and of its hand-derived reverse-mode
adjoint (test helper, CPU only).

This is the *specification* of the adjoint CUDA kernels (fluidgym_b200/csrc, fgb_*_adjoint): the same
table-driven formulas as oracle/table_eval.py, in float64 with exact dense linear solves so that the VJP can
be validated against central finite differences to ~1e-7 (tests/test_adjoint_cpu.py).  The CUDA adjoint
kernels are then compared op by op against these functions on the GPU.

Substep (non-orthogonal path of SIM.py:1431-2002, any number of advect / pressure non-orthogonal iterations):
    (u, p_prev, bvel)  ->  (u_out, p_out)
"""
import numpy as np


class T64:
    """float64 copy of the tables of a CompiledDomain plus a few index helpers."""

    def __init__(self, cd):
        self.N, self.NB = cd.N, cd.NB
        self.nbr = cd.nbr.astype(np.int64)
        self.inner = self.nbr >= 0
        self.nb_safe = np.where(self.inner, self.nbr, 0)
        self.bj = np.where(self.inner, 0, -1 - self.nbr)
        self.comp = (cd.fl_comp & 1).astype(np.int64)
        self.sgn = np.where(cd.fl_comp & 2, -1.0, 1.0)
        for k in ("minv", "det", "Cd", "Wp", "no_gP", "no_gN", "no_wv", "nob_w", "b_minv", "b_det", "b_alpha"):
            setattr(self, k, np.asarray(getattr(cd, k), dtype=np.float64))
        self.b_minv = self.b_minv[:, :cd.NB]
        self.b_det, self.b_alpha = self.b_det[:cd.NB], self.b_alpha[:cd.NB]
        self.no_idx, self.no_face = cd.no_idx.astype(np.int64), cd.no_face.astype(np.int64)
        self.nob_idx = cd.nob_idx.astype(np.int64)
        self.b_face = cd.b_face[:cd.NB].astype(np.int64)
        self.visc = float(cd.visc)
        self.K_no, self.K_nob = cd.K_no, cd.K_nob
        if getattr(cd, "scalar_visc", None) is not None:     # passive scalar (RBC): diffusion table + boundary kinds
            self.Cd_s = np.asarray(cd.Cd_s, dtype=np.float64)
            self.sb_neumann = np.asarray(cd.sb_neumann[:cd.NB]).astype(np.int64)
            self.kappa = float(cd.scalar_visc)
        self.fac = [np.where((~self.inner[2 * d]) | (~self.inner[2 * d + 1]), 1.0, 0.5) for d in range(2)]
        self.cells = np.arange(cd.N)


# ---- elementary pieces ------------------------------------------------------------------------------
def contra(t, v):
    return np.stack([t.det * (t.minv[0] * v[0] + t.minv[1] * v[1]), t.det * (t.minv[2] * v[0] + t.minv[3] * v[1])])


def contra_T(t, Ub):
    """adjoint of contra: v_bar from U_bar"""
    return np.stack([t.det * (t.minv[0] * Ub[0] + t.minv[2] * Ub[1]), t.det * (t.minv[1] * Ub[0] + t.minv[3] * Ub[1])])


def bflux(t, bv):
    ax = t.b_face >> 1
    r0 = np.where(ax == 0, t.b_minv[0], t.b_minv[2])
    r1 = np.where(ax == 0, t.b_minv[1], t.b_minv[3])
    return t.b_det * (r0 * bv[0] + r1 * bv[1]), (t.b_det * r0, t.b_det * r1)


def fluxes(t, v, Fb):
    U = contra(t, v)
    fl = np.zeros((4, t.N))
    for f in range(4):
        velN = t.sgn[f] * U[t.comp[f], t.nb_safe[f]]
        fl[f] = np.where(t.inner[f], 0.5 * (velN + U[f >> 1]), Fb[t.bj[f]])
    return fl


def fluxes_T(t, flb):
    """adjoint of fluxes: (v_bar [2,N], Fb_bar [NB]) from fl_bar [4,N]"""
    Ub = np.zeros((2, t.N))
    Fbb = np.zeros(t.NB)
    for f in range(4):
        g = np.where(t.inner[f], 0.5 * flb[f], 0.0)
        Ub[f >> 1] += g
        np.add.at(Ub, (t.comp[f], t.nb_safe[f]), g * t.sgn[f])
        np.add.at(Fbb, t.bj[f], np.where(t.inner[f], 0.0, flb[f]))
    return contra_T(t, Ub), Fbb


def spmv(t, off, diag, x):
    y = diag * x
    for f in range(4):
        y = y + np.where(t.inner[f], off[f] * x[t.nb_safe[f]], 0.0)
    return y


def dense(t, off, diag):
    M = np.zeros((t.N, t.N))
    M[t.cells, t.cells] = diag
    for f in range(4):
        m = t.inner[f]
        np.add.at(M, (t.cells[m], t.nbr[f][m]), off[f][m])
    return M


def nbr_vals(t, x):
    return np.stack([x] + [np.where(t.inner[f], x[t.nb_safe[f]], x) for f in range(4)])


def nbr_vals_T(t, vb):
    """adjoint of nbr_vals: x_bar from [5,N]"""
    xb = vb[0].copy()
    for f in range(4):
        np.add.at(xb, np.where(t.inner[f], t.nb_safe[f], t.cells), vb[f + 1])
    return xb


# ---- forward substep with a tape ---------------------------------------------------------------------
def substep(t, u, p_prev, bvel, dt, correctors=2, n_adv=1, n_p=1, src=None):
    """n_adv / n_p: advect_non_ortho_steps / pressure_non_ortho_steps (deferred-correction iterations);
    src [2,N]: optional velocity source (buoyancy), added to the predictor right-hand side and to HbyA."""
    tape = {}
    if src is None:
        src = np.zeros((2, t.N))
    Fb, (br0, br1) = bflux(t, bvel)
    fl = fluxes(t, u, Fb)
    sig = np.array([-1.0, 1.0, -1.0, 1.0])
    ff = np.where(t.inner, 0.5 * sig[:, None] * fl, 0.0)
    diag = t.det / dt + t.Cd[0] + ff.sum(0)
    A = diag / t.det
    Coff = np.where(t.inner, (ff + t.Cd[1:]) / t.det, 0.0)
    # boundary sources
    Sb = np.zeros((2, t.N))
    for f in range(4):
        m = ~t.inner[f]
        j = t.bj[f]
        for c in range(2):
            Sb[c] += np.where(m, -bvel[c, j] * (sig[f] * Fb[j]) + bvel[c, j] * 2 * t.visc * t.b_alpha[j], 0.0)
    C = dense(t, Coff, A)
    xs, xprev = [], u
    for _ in range(n_adv):
        NOv = np.zeros((2, t.N))
        for c in range(2):
            for k in range(t.K_no):
                NOv[c] += t.no_wv[k] * xprev[c][t.no_idx[k]]
            for k in range(t.K_nob):
                NOv[c] += t.nob_w[k] * bvel[c][t.nob_idx[k]]
        rhs = (t.det * u / dt + Sb - NOv) / t.det + src
        xprev = np.stack([np.linalg.solve(C, rhs[c]) for c in range(2)])
        xs.append(xprev)
    ustar = xs[-1]
    rA = 1.0 / A
    rAn = nbr_vals(t, rA)
    Pm = np.einsum("ejn,jn->en", t.Wp, rAn)
    P = dense(t, Pm[1:], Pm[0])
    tape.update(Fb=Fb, fl=fl, A=A, Coff=Coff, Sb=Sb, C=C, ustar=ustar, xs=xs, rA=rA, rAn=rAn, Pm=Pm, P=P, cor=[])
    uprev, pprev = ustar, p_prev
    wno = np.stack([t.no_gP[k] * rAn[0] + t.no_gN[k] * rAn[1 + t.no_face[k], t.cells] for k in range(t.K_no)])
    for _ in range(correctors):
        H = np.stack([sum(np.where(t.inner[f], Coff[f] * uprev[c][t.nb_safe[f]], 0.0) for f in range(4)) for c in range(2)])
        hb = rA * (u / dt - H + Sb / t.det + src)
        flh = fluxes(t, hb, Fb)
        sols = []
        for _ps in range(n_p):
            NOp = sum(wno[k] * pprev[t.no_idx[k]] for k in range(t.K_no))
            div = (flh[1] - flh[0]) + (flh[3] - flh[2]) + NOp
            x = np.linalg.solve(P, div)
            p = x - x.mean()
            sols.append(dict(pprev=pprev, x=x, p=p))
            pprev = p
        pv = nbr_vals(t, p)
        pg = np.stack([(pv[2] - pv[1]) * t.fac[0], (pv[4] - pv[3]) * t.fac[1]])
        g = np.stack([pg[0] * t.minv[0] + pg[1] * t.minv[2], pg[0] * t.minv[1] + pg[1] * t.minv[3]])
        unext = hb - rA * g
        tape["cor"].append(dict(uprev=uprev, H=H, hb=hb, sols=sols, g=g, wno=wno))
        uprev = unext
    return uprev, pprev, tape


# ---- reverse pass --------------------------------------------------------------------------------------
def substep_vjp(t, u, p_prev, bvel, dt, tape, u_out_bar, p_out_bar, with_src=False):
    """returns (u_bar, p_prev_bar, bvel_bar) [+ src_bar with_src]"""
    N = t.N
    srcb = np.zeros((2, N))
    sig = np.array([-1.0, 1.0, -1.0, 1.0])
    Coff, A, rA, rAn, Sb, Fb = tape["Coff"], tape["A"], tape["rA"], tape["rAn"], tape["Sb"], tape["Fb"]
    ub = np.zeros((2, N)); bvb = np.zeros((2, t.NB)); Fbb = np.zeros(t.NB)
    Coffb = np.zeros((4, N)); rAb = np.zeros(N); rAnb = np.zeros((5, N)); Sbb = np.zeros((2, N)); Pmb = np.zeros((5, N))
    unb, pb = u_out_bar.copy(), p_out_bar.copy()
    pprevb = np.zeros(N)
    ustarb = np.zeros((2, N)); pprevb_in = np.zeros(N)
    for ci in reversed(range(len(tape["cor"]))):
        cr = tape["cor"][ci]
        # unext = hb - rA * g
        hbb = unb.copy()
        rAb += -(cr["g"] * unb).sum(0)
        gb = -rA * unb
        pgb = np.stack([gb[0] * t.minv[0] + gb[1] * t.minv[1], gb[0] * t.minv[2] + gb[1] * t.minv[3]])
        pvb = np.zeros((5, N))
        pvb[2] += pgb[0] * t.fac[0]; pvb[1] -= pgb[0] * t.fac[0]
        pvb[4] += pgb[1] * t.fac[1]; pvb[3] -= pgb[1] * t.fac[1]
        pb = pb + nbr_vals_T(t, pvb)
        for so in reversed(cr["sols"]):
            # p = x - mean(x) ; x = P^-1 div
            xb = pb - pb.mean()
            lam = np.linalg.solve(tape["P"].T, xb)
            divb = lam
            # P_bar on the pattern: P_ij_bar = -lam_i x_j
            Pmb[0] += -lam * so["x"]
            for f in range(4):
                Pmb[f + 1] += np.where(t.inner[f], -lam * so["x"][t.nb_safe[f]], 0.0)
            # div = flux divergence of hb + NOp(previous pressure iterate)
            flhb = np.stack([-divb, divb, -divb, divb])
            hb_b2, Fbb2 = fluxes_T(t, flhb)
            hbb += hb_b2; Fbb += Fbb2
            pprevb = np.zeros(N)
            for k in range(t.K_no):
                wb = divb * so["pprev"][t.no_idx[k]]
                rAnb[0] += t.no_gP[k] * wb
                np.add.at(rAnb, (1 + t.no_face[k], t.cells), t.no_gN[k] * wb)
                np.add.at(pprevb, t.no_idx[k], cr["wno"][k] * divb)
            pb = pprevb       # the previous iterate enters only through the deferred term of this solve
        # hb = rA (u/dt - H + Sb/det)
        rAb += (hbb * (cr["hb"] / rA)).sum(0)
        inner_b = rA * hbb
        ub += inner_b / dt
        Sbb += inner_b / t.det
        srcb += inner_b
        Hb = -inner_b
        uprevb = np.zeros((2, N))
        for f in range(4):
            m = t.inner[f]
            for c in range(2):
                Coffb[f] += np.where(m, Hb[c] * cr["uprev"][c][t.nb_safe[f]], 0.0)
                np.add.at(uprevb[c], t.nb_safe[f], np.where(m, Hb[c] * Coff[f], 0.0))
        if ci > 0:
            unb = uprevb
        else:
            ustarb, pprevb_in = uprevb, pb
    # P = Wp . rAn
    rAnb += np.einsum("ejn,en->jn", t.Wp, Pmb)
    rAb += nbr_vals_T(t, rAnb)
    Ab = -rAb * rA * rA
    # x_k = C^-1 rhs(u, x_{k-1}),  x_{-1} = u,  ustar = x_{n_adv-1}
    xb = ustarb
    for kk in reversed(range(len(tape["xs"]))):
        xk = tape["xs"][kk]
        rhsb = np.zeros((2, N))
        for c in range(2):
            mu = np.linalg.solve(tape["C"].T, xb[c])
            rhsb[c] = mu
            Ab += -mu * xk[c]
            for f in range(4):
                Coffb[f] += np.where(t.inner[f], -mu * xk[c][t.nb_safe[f]], 0.0)
        # rhs = (det u/dt + Sb - NOv(x_{k-1}))/det
        ub += rhsb / dt
        Sbb += rhsb / t.det
        srcb += rhsb
        NOvb = -rhsb / t.det
        target = ub if kk == 0 else np.zeros((2, N))
        for c in range(2):
            for k in range(t.K_no):
                np.add.at(target[c], t.no_idx[k], t.no_wv[k] * NOvb[c])
            for k in range(t.K_nob):
                np.add.at(bvb[c], t.nob_idx[k], t.nob_w[k] * NOvb[c])
        xb = target
    # Sb(bvel, Fb)
    for f in range(4):
        m = ~t.inner[f]
        j = t.bj[f]
        for c in range(2):
            np.add.at(bvb[c], j, np.where(m, Sbb[c] * (-(sig[f] * Fb[j]) + 2 * t.visc * t.b_alpha[j]), 0.0))
            np.add.at(Fbb, j, np.where(m, Sbb[c] * (-bvel[c, j] * sig[f]), 0.0))
    # A = diag/det, Coff = (ff + Cd)/det, diag = det/dt + Cd0 + sum ff
    diagb = Ab / t.det
    ffb = np.where(t.inner, Coffb / t.det + diagb[None, :], 0.0)
    flb = 0.5 * sig[:, None] * ffb
    ub2, Fbb3 = fluxes_T(t, flb)
    ub += ub2; Fbb += Fbb3
    # Fb(bvel)
    _, (br0, br1) = bflux(t, bvel)
    bvb[0] += br0 * Fbb
    bvb[1] += br1 * Fbb
    if with_src:
        return ub, pprevb_in, bvb, srcb
    return ub, pprevb_in, bvb


# ---- passive scalar + buoyancy (RBC substep: SIM.py:1471-1657, rbc_env_base.py:280-304) -------------------------
SIG = np.array([-1.0, 1.0, -1.0, 1.0])


def scalar_step(t, u, bvel, T, sbval, dt):
    """T_new = C_s(u)^-1 rhs_s(T, bvel, sbval) (k_setup_scalar + BiCGStab), with a tape."""
    Fb, _ = bflux(t, bvel)
    fl = fluxes(t, u, Fb)
    ff = np.where(t.inner, 0.5 * SIG[:, None] * fl, 0.0)
    diag = t.det / dt + t.Cd_s[0] + ff.sum(0)
    As = diag / t.det
    Coffs = np.where(t.inner, (ff + t.Cd_s[1:]) / t.det, 0.0)
    r = t.det * T / dt
    for f in range(4):
        m = ~t.inner[f]
        j = t.bj[f]
        dif = t.kappa * np.where(t.sb_neumann[j] == 0, 2.0 * t.b_alpha[j], 1.0)
        r = r + np.where(m, sbval[j] * (-(SIG[f] * Fb[j]) + dif), 0.0)
    Cs = dense(t, Coffs, As)
    Tn = np.linalg.solve(Cs, r / t.det)
    return Tn, dict(Cs=Cs, Tn=Tn, Fb=Fb)


def scalar_step_vjp(t, u, bvel, T, sbval, dt, tape, Tn_bar):
    """returns (u_bar, bvel_bar, T_bar, sbval_bar)"""
    N = t.N
    Tn, Fb = tape["Tn"], tape["Fb"]
    lam = np.linalg.solve(tape["Cs"].T, Tn_bar)
    Asb = -lam * Tn
    Coffsb = np.stack([np.where(t.inner[f], -lam * Tn[t.nb_safe[f]], 0.0) for f in range(4)])
    rb = lam / t.det
    Tb = rb * t.det / dt
    sbb = np.zeros(t.NB); Fbb = np.zeros(t.NB)
    for f in range(4):
        m = ~t.inner[f]
        j = t.bj[f]
        dif = t.kappa * np.where(t.sb_neumann[j] == 0, 2.0 * t.b_alpha[j], 1.0)
        np.add.at(sbb, j, np.where(m, rb * (-(SIG[f] * Fb[j]) + dif), 0.0))
        np.add.at(Fbb, j, np.where(m, -rb * sbval[j] * SIG[f], 0.0))
    diagb = Asb / t.det
    ffb = np.where(t.inner, Coffsb / t.det + diagb[None, :], 0.0)
    ub, Fbb2 = fluxes_T(t, 0.5 * SIG[:, None] * ffb)
    Fbb += Fbb2
    _, (br0, br1) = bflux(t, bvel)
    return ub, np.stack([br0 * Fbb, br1 * Fbb]), Tb, sbb


def substep_scalar(t, u, p_prev, bvel, T, sbval, dt, beta, correctors=2):
    """RBC substep: scalar transport with the OLD velocity, buoyancy source from the NEW temperature, PISO substep."""
    Tn, stape = scalar_step(t, u, bvel, T, sbval, dt)
    src = np.stack([np.zeros(t.N), beta * Tn])
    uo, po, tape = substep(t, u, p_prev, bvel, dt, correctors=correctors, src=src)
    return uo, po, Tn, dict(piso=tape, scalar=stape)


def substep_scalar_vjp(t, u, p_prev, bvel, T, sbval, dt, beta, tape, u_out_bar, p_out_bar, T_out_bar):
    """returns (u_bar, p_prev_bar, bvel_bar, T_bar, sbval_bar)"""
    ub, pb, bvb, srcb = substep_vjp(t, u, p_prev, bvel, dt, tape["piso"], u_out_bar, p_out_bar, with_src=True)
    Tnb = T_out_bar + beta * srcb[1]
    ub2, bvb2, Tb, sbb = scalar_step_vjp(t, u, bvel, T, sbval, dt, tape["scalar"], Tnb)
    return ub + ub2, pb, bvb + bvb2, Tb, sbb
