// D = 3 operators on z-EXTRUDED multi-block domains (CylinderJet3D, Airfoil3D: the 2-D multi-block grid repeated over nz
// uniform, periodic z planes; envs/cylinder/grid.py:298, shapes.py:641-676; reference: the K.cu kernels with DIMS = 3).
//
// STATUS: operator layer.  The per-cell functions below are __host__ __device__ so that tests/test_extruded_host.py can
// execute exactly this code on the CPU (through tests/cpu_harness/extruded_host.cu) and compare it with the numpy
// specification oracle/extruded_eval.py, which is pinned to an op trace of the unmodified reference on CylinderJet3D-easy
// (tests/golden/cyl3d_substep*.npz).  The launch glue at the end of this file has NOT run on a GPU yet
// (DESIGN.md section 9): envs/cylinder3d.py (CylinderJet3D) is built on it and verified on the CPU
// through the same cell code (tests/test_cylinder3d_cpu.py); tools/extruded_check.py is its first GPU run.
//
// The metric tensor of an extruded cell is block diagonal, M3 = diag(M2, hz): every in-plane coefficient of a row / det3
// equals the 2-D one from the compiled tables (fgb_tables), the z faces add  -+1/4 (u_z,P + u_z,N) / hz - nu / hz^2  off
// the diagonal; all non-orthogonal (corner) terms stay in the x-y plane because alpha^{xz} = alpha^{yz} = 0.
//
// Layout: cell g3 = k * N2 + g (plane k, 2-D cell g); fields [B][3][N3]; boundary velocities [B][3][nz][NB2]; matrices
// ELL(7) on a 6-face neighbour table (faces 0..3 in plane, 4 = -z, 5 = +z) so that the cooperative Krylov kernels of
// ortho3_b200.cuh (k3_bicgstab, k3_cg) solve them unchanged.
#pragma once

#ifndef X3_HD
#define X3_HD __host__ __device__ __forceinline__
#endif

struct X3Tab {
    fgb_tables t;     // in-plane tables of the compiled 2-D domain
    int nz;           // z planes (periodic)
    float hz;         // plane spacing
};

X3_HD float x3_contra(const fgb_tables &t, int k, int g, float u, float v) {          // det2 * (Minv[k] . (u, v))   (K.cu:495-510)
    const int N = t.N;
    return t.det[g] * (t.minv[(2 * k) * N + g] * u + t.minv[(2 * k + 1) * N + g] * v);
}
X3_HD float x3_bflux(const fgb_tables &t, int j, int ax, float bu, float bv) {
    const int NB = t.NB;
    return t.b_det[j] * (t.b_minv[(2 * ax) * NB + j] * bu + t.b_minv[(2 * ax + 1) * NB + j] * bv);
}
// in-plane face fluxes / hz of plane-local fields ux, uy [N2] (K.cu:1567-1645)
X3_HD void x3_face_fluxes(const fgb_tables &t, int g, const float *ux, const float *uy, const float *bx, const float *by,
                          const int nb[4], float fl[4]) {
    const int N = t.N;
    const float Uc[2] = {x3_contra(t, 0, g, ux[g], uy[g]), x3_contra(t, 1, g, ux[g], uy[g])};
    for (int f = 0; f < 4; ++f) {
        if (nb[f] >= 0) {
            const int fc = t.fl_comp[f * N + g];
            float velN = x3_contra(t, fc & 1, nb[f], ux[nb[f]], uy[nb[f]]);
            if (fc & 2) velN = -velN;
            fl[f] = (velN + Uc[f >> 1]) * 0.5f;
        } else {
            const int j = -1 - nb[f];
            fl[f] = x3_bflux(t, j, f >> 1, bx[j], by[j]);
        }
    }
}
// Dirichlet boundary advection + diffusion source of component values bc [NB2] (K.cu:4321-4380), before / det
X3_HD float x3_boundary_source(const fgb_tables &t, const float *bx, const float *by, const float *bc, const int nb[4]) {
    float S = 0.f;
    for (int f = 0; f < 4; ++f)
        if (nb[f] < 0) {
            const int j = -1 - nb[f];
            const float flux = x3_bflux(t, j, f >> 1, bx[j], by[j]) * ((f & 1) ? 1.f : -1.f);
            S -= bc[j] * flux;
            S += bc[j] * (t.viscosity * 2.f * t.b_alpha[j]);
        }
    return S;
}

// SetupAdvectionMatrix + SetupAdvectionVelocity (K.cu:3617-3880, 4296-4400) for cell (k, g) of one environment.
// u, ures: [3][N3]; bvel: [3][nz][NB2]; outputs coff [6][N3], A [N3], rhs [3][N3]
X3_HD void x3_setup_advection_cell(const X3Tab &x, int k, int g, const float *u, const float *ures, const float *bvel, float dt,
                                   float *coff, float *A, float *rhs, int with_matrix) {
    const fgb_tables &t = x.t;
    const int N = t.N, NB = t.NB, nz = x.nz, N3 = N * nz, g3 = k * N + g;
    const float det = t.det[g], hz = x.hz;
    const float *ux = u + (size_t)k * N, *uy = u + (size_t)N3 + (size_t)k * N, *uz = u + 2 * (size_t)N3;
    const float *bx = bvel + (size_t)k * NB, *by = bvel + (size_t)nz * NB + (size_t)k * NB;
    int nb[4];
    for (int f = 0; f < 4; ++f) nb[f] = t.nbr[f * N + g];
    if (with_matrix) {
        float fl[4];
        x3_face_fluxes(t, g, ux, uy, bx, by, nb, fl);
        float diag = det / dt + t.Cd[g];
        for (int f = 0; f < 4; ++f) {
            float o = 0.f;
            if (nb[f] >= 0) {
                const float ff = ((f & 1) ? 0.5f : -0.5f) * fl[f];
                diag += ff;
                o = (ff + t.Cd[(f + 1) * N + g]) / det;
            }
            coff[(size_t)f * N3 + g3] = o;
        }
        const int kl = (k + nz - 1) % nz, ku = (k + 1) % nz;
        const float uzP = uz[g3], Fzm = 0.5f * (uzP + uz[(size_t)kl * N + g]), Fzp = 0.5f * (uzP + uz[(size_t)ku * N + g]);
        const float dz = t.viscosity / (hz * hz);
        coff[(size_t)4 * N3 + g3] = (-0.5f * Fzm) / hz - dz;
        coff[(size_t)5 * N3 + g3] = (0.5f * Fzp) / hz - dz;
        A[g3] = diag / det + 2.f * dz + 0.5f * (Fzp - Fzm) / hz;
    }
    for (int c = 0; c < 3; ++c) {
        const float *bc = bvel + ((size_t)c * nz + k) * NB;
        const float *ur = ures + (size_t)c * N3 + (size_t)k * N;
        const float S = x3_boundary_source(t, bx, by, bc, nb);
        float no = 0.f;
        for (int q = 0; q < t.K_no; ++q) {
            const float w = t.no_wv[q * N + g];
            if (w != 0.f) no += w * ur[t.no_idx[q * N + g]];
        }
        for (int q = 0; q < t.K_nob; ++q) {
            const float w = t.nob_w[q * N + g];
            if (w != 0.f) no += w * bc[t.nob_idx[q * N + g]];
        }
        rhs[(size_t)c * N3 + g3] = (det * u[(size_t)c * N3 + g3] / dt + S - no) / det;
    }
}

// SetupPressureMatrix (K.cu:4812-4978), NOT divided by det: in-plane  hz * sum_j Wp[e][j] (1/A)_j , z faces 1/2 (det2 / hz) (1/A_P + 1/A_N)
X3_HD void x3_pressure_matrix_cell(const X3Tab &x, int k, int g, const float *A, float *poff, float *pdiag) {
    const fgb_tables &t = x.t;
    const int N = t.N, nz = x.nz, N3 = N * nz, g3 = k * N + g;
    const float *a = A + (size_t)k * N;
    float rA[5];
    rA[0] = 1.0f / a[g];
    for (int f = 0; f < 4; ++f) { const int nb = t.nbr[f * N + g]; rA[f + 1] = nb >= 0 ? 1.0f / a[nb] : rA[0]; }
    float P[5];
    for (int e = 0; e < 5; ++e) {
        float s = 0.f;
        for (int j = 0; j < 5; ++j) s += t.Wp[(5 * e + j) * N + g] * rA[j];
        P[e] = s * x.hz;
    }
    const int kl = (k + nz - 1) % nz, ku = (k + 1) % nz;
    const float az = t.det[g] / x.hz;
    const float pl = 0.5f * az * (rA[0] + 1.0f / A[(size_t)kl * N + g]), pu = 0.5f * az * (rA[0] + 1.0f / A[(size_t)ku * N + g]);
    for (int f = 0; f < 4; ++f) poff[(size_t)f * N3 + g3] = P[f + 1];
    poff[(size_t)4 * N3 + g3] = pl;
    poff[(size_t)5 * N3 + g3] = pu;
    pdiag[g3] = P[0] - pl - pu;
}

// PISO_build_pressure_rhs (K.cu:5136-5255): HbyA = (u/dt - sum_nb C_nb u*_nb + S_b/det) / A for the three components
X3_HD void x3_hbya_cell(const X3Tab &x, int k, int g, const float *u, const float *ures, const float *bvel, const float *coff,
                        const float *A, float dt, float *hb) {
    const fgb_tables &t = x.t;
    const int N = t.N, NB = t.NB, nz = x.nz, N3 = N * nz, g3 = k * N + g;
    const float *bx = bvel + (size_t)k * NB, *by = bvel + (size_t)nz * NB + (size_t)k * NB;
    const int kl = (k + nz - 1) % nz, ku = (k + 1) % nz;
    int nb[4];
    for (int f = 0; f < 4; ++f) nb[f] = t.nbr[f * N + g];
    const float rD = 1.0f / A[g3], det = t.det[g];
    for (int c = 0; c < 3; ++c) {
        const float *ur = ures + (size_t)c * N3;
        float H = 0.f;
        for (int f = 0; f < 4; ++f)
            if (nb[f] >= 0) H += coff[(size_t)f * N3 + g3] * ur[(size_t)k * N + nb[f]];
        H += coff[(size_t)4 * N3 + g3] * ur[(size_t)kl * N + g] + coff[(size_t)5 * N3 + g3] * ur[(size_t)ku * N + g];
        const float S = x3_boundary_source(t, bx, by, bvel + ((size_t)c * nz + k) * NB, nb);
        hb[(size_t)c * N3 + g3] = rD * (u[(size_t)c * N3 + g3] / dt - H + S / det);
    }
}

// divergence of the face fluxes of HbyA (K.cu:5389-5434) + deferred non-orthogonal pressure term (K.cu:5470-5492)
X3_HD void x3_divergence_cell(const X3Tab &x, int k, int g, const float *hb, const float *bvel, const float *pprev /* or null */,
                              const float *A, float *div) {
    const fgb_tables &t = x.t;
    const int N = t.N, NB = t.NB, nz = x.nz, N3 = N * nz, g3 = k * N + g;
    const float *hx = hb + (size_t)k * N, *hy = hb + (size_t)N3 + (size_t)k * N, *hzc = hb + 2 * (size_t)N3;
    const float *bx = bvel + (size_t)k * NB, *by = bvel + (size_t)nz * NB + (size_t)k * NB;
    int nb[4];
    for (int f = 0; f < 4; ++f) nb[f] = t.nbr[f * N + g];
    float fl[4];
    x3_face_fluxes(t, g, hx, hy, bx, by, nb, fl);
    float d = (fl[1] - fl[0]) + (fl[3] - fl[2]);
    if (pprev) {
        const float *a = A + (size_t)k * N, *pp = pprev + (size_t)k * N;
        float rA[5];
        rA[0] = 1.0f / a[g];
        for (int f = 0; f < 4; ++f) rA[f + 1] = nb[f] >= 0 ? 1.0f / a[nb[f]] : rA[0];
        float S = 0.f;
        for (int q = 0; q < t.K_no; ++q) {
            const float gP = t.no_gP[q * N + g], gN = t.no_gN[q * N + g];
            if (gP != 0.f || gN != 0.f) S += (gP * rA[0] + gN * rA[1 + t.no_face[q * N + g]]) * pp[t.no_idx[q * N + g]];
        }
        d += S;
    }
    const int kl = (k + nz - 1) % nz, ku = (k + 1) % nz;
    const float hP = hzc[g3];
    div[g3] = d * x.hz + t.det[g] * (0.5f * (hP + hzc[(size_t)ku * N + g]) - 0.5f * (hP + hzc[(size_t)kl * N + g]));
}

// PISO_update_velocity (K.cu:816-849, 5962-5995)
X3_HD void x3_correct_cell(const X3Tab &x, int k, int g, const float *hb, const float *p, const float *A, float *uout) {
    const fgb_tables &t = x.t;
    const int N = t.N, nz = x.nz, N3 = N * nz, g3 = k * N + g;
    const float *pk = p + (size_t)k * N;
    const float pc = pk[g];
    float pg[2];
    for (int d = 0; d < 2; ++d) {
        const int nl = t.nbr[(2 * d) * N + g], nu = t.nbr[(2 * d + 1) * N + g];
        const float fac = (nl < 0 || nu < 0) ? 1.0f : 0.5f;
        pg[d] = ((nu >= 0 ? pk[nu] : pc) - (nl >= 0 ? pk[nl] : pc)) * fac;
    }
    const float gx = pg[0] * t.minv[g] + pg[1] * t.minv[2 * N + g];
    const float gy = pg[0] * t.minv[N + g] + pg[1] * t.minv[3 * N + g];
    const int kl = (k + nz - 1) % nz, ku = (k + 1) % nz;
    const float gz = 0.5f * (p[(size_t)ku * N + g] - p[(size_t)kl * N + g]) / x.hz;
    const float rD = 1.0f / A[g3];
    uout[g3] = hb[g3] - rD * gx;
    uout[(size_t)N3 + g3] = hb[(size_t)N3 + g3] - rD * gy;
    uout[2 * (size_t)N3 + g3] = hb[2 * (size_t)N3 + g3] - rD * gz;
}

#ifdef __CUDACC__
#ifndef X3_HOST_ONLY
// ------------------------------------------------------------------------------------------------------------------
// launch glue (one thread per (cell, plane, environment)); matrices / vectors live in the workspace of an fgb_ortho3
// handle created on the 6-face neighbour table of the extruded domain, so fgb_ortho3_solve_advection / _solve_pressure
// run the Krylov iterations.  NOT yet run on a GPU (see the header of this file).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kx3_setup_advection(X3Tab x, const float *U, const float *Ures, const float *Bvel, const float *dtv,
                                                          float *Coff, float *A, float *Rhs, int with_matrix) {
    const int b = blockIdx.z, k = blockIdx.y, g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= x.t.N) return;
    const size_t N3 = (size_t)x.t.N * x.nz;
    x3_setup_advection_cell(x, k, g, U + b * 3 * N3, Ures + b * 3 * N3, Bvel + (size_t)b * 3 * x.nz * x.t.NB, dtv[b], Coff + b * 6 * N3,
                            A + b * N3, Rhs + b * 3 * N3, with_matrix);
}
__global__ void __launch_bounds__(256) kx3_pressure_matrix(X3Tab x, const float *A, float *Poff, float *Pdiag) {
    const int b = blockIdx.z, k = blockIdx.y, g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= x.t.N) return;
    const size_t N3 = (size_t)x.t.N * x.nz;
    x3_pressure_matrix_cell(x, k, g, A + b * N3, Poff + b * 6 * N3, Pdiag + b * N3);
}
__global__ void __launch_bounds__(256) kx3_hbya(X3Tab x, const float *U, const float *Ures, const float *Bvel, const float *Coff, const float *A,
                                               const float *dtv, float *Hb) {
    const int b = blockIdx.z, k = blockIdx.y, g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= x.t.N) return;
    const size_t N3 = (size_t)x.t.N * x.nz;
    x3_hbya_cell(x, k, g, U + b * 3 * N3, Ures + b * 3 * N3, Bvel + (size_t)b * 3 * x.nz * x.t.NB, Coff + b * 6 * N3, A + b * N3, dtv[b], Hb + b * 3 * N3);
}
__global__ void __launch_bounds__(256) kx3_divergence(X3Tab x, const float *Hb, const float *Bvel, const float *Pprev, const float *A, float *Div) {
    const int b = blockIdx.z, k = blockIdx.y, g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= x.t.N) return;
    const size_t N3 = (size_t)x.t.N * x.nz;
    x3_divergence_cell(x, k, g, Hb + b * 3 * N3, Bvel + (size_t)b * 3 * x.nz * x.t.NB, Pprev ? Pprev + b * N3 : nullptr, A + b * N3, Div + b * N3);
}
__global__ void __launch_bounds__(256) kx3_correct(X3Tab x, const float *Hb, const float *P, const float *A, float *Uout) {
    const int b = blockIdx.z, k = blockIdx.y, g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= x.t.N) return;
    const size_t N3 = (size_t)x.t.N * x.nz;
    x3_correct_cell(x, k, g, Hb + b * 3 * N3, P + b * N3, A + b * N3, Uout + b * 3 * N3);
}

// Simulation._PISO_split_step on an extruded domain (SIM.py:1431-2002, non-orthogonal path): b = fgb_ortho3 handle created on
// the 6-face table (N = nz * N2); opt.adv_nonortho_steps / p_nonortho_steps deferred-correction iterations.
extern "C" int fgb_extruded3_piso_substep(fgb_ortho3 *b, const fgb_extruded3_tables *xt, float *u, float *p, const float *bvel, const float *dt,
                                          fgb_stream_t s) {
    if (!b || !xt || !u || !p || !bvel || !dt) return set_err(FGB_E_ARG, "fgb_extruded3_piso_substep: null argument");
    if (b->t.N != xt->plane.N * xt->nz || b->slab.on) return set_err(FGB_E_ARG, "fgb_extruded3_piso_substep: handle / tables mismatch");
    X3Tab x; x.t = xt->plane; x.nz = xt->nz; x.hz = xt->hz;
    cudaStream_t st = STREAM(s);
    const dim3 grid((unsigned)((x.t.N + 255) / 256), (unsigned)x.nz, (unsigned)b->B);
    const fgb_options &o = b->opt;
    int rc;
    for (int ns = 0; ns < o.adv_nonortho_steps; ++ns) {
        b->launches++;
        kx3_setup_advection<<<grid, 256, 0, st>>>(x, u, ns == 0 ? u : b->ures, bvel, dt, b->Coff, b->A, b->rhs, ns == 0);
        LAUNCH_CHECK("kx3_setup_advection");
        if ((rc = fgb_ortho3_solve_advection(b, ns == 0, nullptr, s))) return rc;
    }
    for (int cs = 0; cs < o.corrector_steps; ++cs) {
        b->launches += 2;
        if (cs == 0) { kx3_pressure_matrix<<<grid, 256, 0, st>>>(x, b->A, b->Poff, b->Pdiag); LAUNCH_CHECK("kx3_pressure_matrix"); }
        kx3_hbya<<<grid, 256, 0, st>>>(x, u, b->ures, bvel, b->Coff, b->A, dt, b->hbya);
        LAUNCH_CHECK("kx3_hbya");
        for (int ps = 0; ps < o.p_nonortho_steps; ++ps) {
            b->launches++;
            kx3_divergence<<<grid, 256, 0, st>>>(x, b->hbya, bvel, p, b->A, b->div);
            LAUNCH_CHECK("kx3_divergence");
            if ((rc = fgb_ortho3_solve_pressure(b, p, ps == 0, 100, o.max_iter, cs * o.p_nonortho_steps + ps, nullptr, s))) return rc;
        }
        b->launches++;
        kx3_correct<<<grid, 256, 0, st>>>(x, b->hbya, p, b->A, b->ures);
        LAUNCH_CHECK("kx3_correct");
    }
    cudaError_t ce = cudaMemcpyAsync(u, b->ures, (size_t)b->B * 3 * b->t.N * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (ce != cudaSuccess) return set_err(FGB_E_CUDA, "fgb_extruded3_piso_substep: copy", ce);
    return FGB_OK;
}

// Simulation.make_divergence_free on an extruded domain (SIM.py:1320-1430): A = 1, the velocity itself is the pressure right-hand
// side vector; p_nonortho_steps x [divergence + deferred non-orthogonal term of the current pressure, CG (zero start in the first
// iteration, then restarted from the previous result), mean removal], one velocity correction.  The "PRE" hook (outflow update
// with time step 1) is the caller's.
extern "C" int fgb_extruded3_make_divergence_free(fgb_ortho3 *b, const fgb_extruded3_tables *xt, float *u, float *p, const float *bvel,
                                                  int max_iter, fgb_stream_t s) {
    if (!b || !xt || !u || !p || !bvel) return set_err(FGB_E_ARG, "fgb_extruded3_make_divergence_free: null argument");
    if (b->t.N != xt->plane.N * xt->nz || b->slab.on) return set_err(FGB_E_ARG, "fgb_extruded3_make_divergence_free: handle / tables mismatch");
    X3Tab x; x.t = xt->plane; x.nz = xt->nz; x.hz = xt->hz;
    cudaStream_t st = STREAM(s);
    const dim3 grid((unsigned)((x.t.N + 255) / 256), (unsigned)x.nz, (unsigned)b->B);
    const size_t BN = (size_t)b->B * b->t.N;
    int rc;
    b->launches += 2;
    k_fill<<<(unsigned)((BN + 255) / 256), 256, 0, st>>>(b->A, 1.0f, BN);
    LAUNCH_CHECK("k_fill");
    kx3_pressure_matrix<<<grid, 256, 0, st>>>(x, b->A, b->Poff, b->Pdiag);
    LAUNCH_CHECK("kx3_pressure_matrix");
    for (int ps = 0; ps < b->opt.p_nonortho_steps; ++ps) {
        b->launches++;
        kx3_divergence<<<grid, 256, 0, st>>>(x, u, bvel, p, b->A, b->div);
        LAUNCH_CHECK("kx3_divergence");
        if ((rc = fgb_ortho3_solve_pressure(b, p, ps == 0, 0 /* no residual reset here, SIM.py:1387-1396 */, max_iter > 0 ? max_iter : b->opt.max_iter, ps, nullptr, s))) return rc;
    }
    b->launches++;
    kx3_correct<<<grid, 256, 0, st>>>(x, u, p, b->A, b->ures);
    LAUNCH_CHECK("kx3_correct");
    cudaError_t ce = cudaMemcpyAsync(u, b->ures, 3 * BN * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (ce != cudaSuccess) return set_err(FGB_E_CUDA, "fgb_extruded3_make_divergence_free: copy", ce);
    return FGB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Boundary hooks of the extruded environments as kernels (OPT-IN: FGB_X3_HOOKS=cuda in extruded3d.py; default = the torch
// expressions of ExtrudedStepping, which are what the CPU tests pin to the reference).  Same formulas, statement by statement:
// balance_boundary_fluxes (SIM.py:188-224), update_advective_boundaries (SIM.py:228-393), Domain.getMaxVelocity (DS.cpp:1580-1612).
// One CTA per environment for the boundary kernels (a few thousand faces), flux sums accumulated in double.  NOT yet run on a GPU.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void x3_block_sum2(double &a, double &b, double *sm /* [66] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) { sm[2 * warp] = a; sm[2 * warp + 1] = b; }
    __syncthreads();
    if (warp == 0) {
        double x = lane < nw ? sm[2 * lane] : 0.0, y = lane < nw ? sm[2 * lane + 1] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { x += __shfl_xor_sync(0xffffffffu, x, o); y += __shfl_xor_sync(0xffffffffu, y, o); }
        if (lane == 0) { sm[64] = x; sm[65] = y; }
    }
    __syncthreads();
    a = sm[64]; b = sm[65];
    __syncthreads();
}
// scale all components of the free faces by -(flux through the other prescribed faces) / (flux through the free faces)
__device__ void x3_balance(int NB, int nz, float hz, float *bv /* [3][nz][NB] */, const float *__restrict__ fw /* [2][NB] */,
                           const int8_t *__restrict__ free_mask, float tol, double *sm) {
    const int n = nz * NB;
    double fixed = 0.0, var = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int j = i % NB;
        const float fl = (bv[i] * fw[j] + bv[n + i] * fw[NB + j]) * hz;
        if (free_mask[j]) var += (double)fl; else fixed += (double)fl;
    }
    x3_block_sum2(fixed, var, sm);
    if (!(fabs(fixed + var) <= (double)tol * 0.01)) {
        const float sc = (float)(-fixed / var);
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            if (free_mask[i % NB]) { bv[i] *= sc; bv[n + i] *= sc; bv[2 * n + i] *= sc; }
    }
    __syncthreads();
}
__global__ void __launch_bounds__(512) kx3_balance_fluxes(int NB, int nz, float hz, float *Bvel, const float *__restrict__ fw,
                                                          const int8_t *__restrict__ free_mask, float tol) {
    __shared__ double sm[66];
    x3_balance(NB, nz, hz, Bvel + (size_t)blockIdx.x * 3 * nz * NB, fw, free_mask, tol, sm);
}
__global__ void __launch_bounds__(512) kx3_update_outflow(int N2, int NB, int nz, float hz, const float *__restrict__ U, float *Bvel,
                                                          const float *__restrict__ dtv, const float *__restrict__ fw,
                                                          const int8_t *__restrict__ out_mask, int n_out, const int32_t *__restrict__ out_face,
                                                          const int32_t *__restrict__ out_cell, const float *__restrict__ out_adv, float tol) {
    __shared__ double sm[66];
    const int b = blockIdx.x, n = nz * NB;
    const size_t N3 = (size_t)N2 * nz;
    const float *u = U + (size_t)b * 3 * N3;
    float *bv = Bvel + (size_t)b * 3 * n;
    const float dt = dtv[b];
    for (int i = threadIdx.x; i < nz * n_out; i += blockDim.x) {
        const int k = i / n_out, q = i % n_out, j = out_face[q], c = out_cell[q];
        const float w = 1.0f - 1.0f / (1.0f + 2.0f * dt * out_adv[q]);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float bo = bv[(size_t)d * n + k * NB + j];
            bv[(size_t)d * n + k * NB + j] = bo - w * (bo - u[(size_t)d * N3 + (size_t)k * N2 + c]);
        }
    }
    __syncthreads();
    x3_balance(NB, nz, hz, bv, fw, out_mask, tol, sm);
}
__global__ void __launch_bounds__(256) kx3_max_velocity(X3Tab x, const float *__restrict__ U, const float *__restrict__ Bvel, float *__restrict__ maxvel) {
    __shared__ float smf[33];
    const int b = blockIdx.y, N2 = x.t.N, NB = x.t.NB, nz = x.nz;
    const size_t N3 = (size_t)N2 * nz;
    const float *u = U + (size_t)b * 3 * N3, *bv = Bvel + (size_t)b * 3 * nz * NB;
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N3; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % N2);
        const float ux = u[i], uy = u[N3 + i];
        m = fmaxf(m, fabsf(x.t.minv[g] * ux + x.t.minv[N2 + g] * uy));
        m = fmaxf(m, fabsf(x.t.minv[2 * N2 + g] * ux + x.t.minv[3 * N2 + g] * uy));
        m = fmaxf(m, fabsf(u[2 * N3 + i]) / x.hz);
    }
    const int nb = nz * NB;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += gridDim.x * blockDim.x) {
        const int j = i % NB;
        const float bx = bv[i], by = bv[nb + i];
        m = fmaxf(m, fabsf(x.t.b_minv[j] * bx + x.t.b_minv[NB + j] * by));
        m = fmaxf(m, fabsf(x.t.b_minv[2 * NB + j] * bx + x.t.b_minv[3 * NB + j] * by));
        m = fmaxf(m, fabsf(bv[2 * nb + i]) / x.hz);
    }
    const float mv = block_reduce_max(m, smf);
    if (threadIdx.x == 0) atomicMax((int *)&maxvel[b], __float_as_int(mv));
}

// per-plane wall forces (forces.py:278-377 = the 2-D wall traction of every plane times the plane spacing; same statements as
// k_wall_forces with the extruded strides): out[B][nz][2] (drag, lift coefficient contributions), overwritten
__global__ void __launch_bounds__(128) kx3_wall_forces(fgb_wall w, float visc, int N2, int NB, int nz, float hz, const float *__restrict__ U,
                                                        const float *__restrict__ P, const float *__restrict__ Bvel, float *__restrict__ out) {
    __shared__ double red[32 * 2 + 2];
    const int k = blockIdx.x, b = blockIdx.y;
    const size_t N3 = (size_t)N2 * nz, nb3 = (size_t)nz * NB;
    const float *u = U + (size_t)b * 3 * N3 + (size_t)k * N2, *p = P + (size_t)b * N3 + (size_t)k * N2;
    const float *bv = Bvel + (size_t)b * 3 * nb3 + (size_t)k * NB;
    const int n = w.n_wall;
    float f[2] = {0.f, 0.f};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = w.cell[i], j = w.bface[i];
        const int il = w.cell[(i + 1) % n], ir = w.cell[(i + n - 1) % n];
        const float nx = w.normal[i], ny = w.normal[n + i];
        const float tx = ny, ty = -nx;
        const float du_dn = (u[c] - bv[j]) / w.dist[i], dv_dn = (u[N3 + c] - bv[nb3 + j]) / w.dist[i];
        const float du_dt = (u[ir] - u[il]) / (2.f * w.tlen[i]), dv_dt = (u[N3 + ir] - u[N3 + il]) / (2.f * w.tlen[i]);
        const float du_dx = du_dn * nx + du_dt * tx, du_dy = du_dn * ny + du_dt * ty;
        const float dv_dx = dv_dn * nx + dv_dt * tx, dv_dy = dv_dn * ny + dv_dt * ty;
        const float sxx = 2.f * visc * du_dx - p[c], syy = 2.f * visc * dv_dy - p[c];
        const float sxy = 2.f * visc * (0.5f * (du_dy + dv_dx));
        f[0] += (sxx * nx + sxy * ny) * w.flen[i];
        f[1] += (sxy * nx + syy * ny) * w.flen[i];
    }
    block_reduce_sum<2>(f, red);
    if (threadIdx.x == 0) { out[((size_t)b * nz + k) * 2] = f[0] * w.scale * hz; out[((size_t)b * nz + k) * 2 + 1] = f[1] * w.scale * hz; }
}
extern "C" int fgb_extruded3_wall_forces(const fgb_extruded3_tables *xt, int32_t B, const fgb_wall *w, const float *u, const float *p,
                                         const float *bvel, float *out, fgb_stream_t s) {
    if (!xt || !w || !u || !p || !bvel || !out || B <= 0) return set_err(FGB_E_ARG, "fgb_extruded3_wall_forces: bad argument");
    kx3_wall_forces<<<dim3((unsigned)xt->nz, (unsigned)B), 128, 0, STREAM(s)>>>(*w, xt->plane.viscosity, xt->plane.N, xt->plane.NB, xt->nz, xt->hz,
                                                                               u, p, bvel, out);
    LAUNCH_CHECK("kx3_wall_forces");
    return FGB_OK;
}
// actuation (jet_cylinder_env_3d.py:399-424, airfoil_env_3d.py:383-407): bvel[0..1][k][face q] = sum_j amp[b][k][j] * templ[j][0..1][q],
// spanwise component 0, then the flux balance over the free faces (jets + outflow)
__global__ void __launch_bounds__(512) kx3_apply_jets(int NB, int nz, float hz, float *Bvel, const float *__restrict__ amp, int J,
                                                      const int32_t *__restrict__ jet_face, const float *__restrict__ templ, int nf,
                                                      const float *__restrict__ fw, const int8_t *__restrict__ free_mask, float tol) {
    __shared__ double sm[66];
    const int b = blockIdx.x, n = nz * NB;
    float *bv = Bvel + (size_t)b * 3 * n;
    const float *a = amp + (size_t)b * nz * J;
    for (int i = threadIdx.x; i < nz * nf; i += blockDim.x) {
        const int k = i / nf, q = i % nf, j = jet_face[q];
        float v0 = 0.f, v1 = 0.f;
        for (int jj = 0; jj < J; ++jj) { const float am = a[k * J + jj]; v0 += am * templ[(jj * 2) * nf + q]; v1 += am * templ[(jj * 2 + 1) * nf + q]; }
        bv[k * NB + j] = v0; bv[n + k * NB + j] = v1; bv[2 * n + k * NB + j] = 0.f;
    }
    __syncthreads();
    x3_balance(NB, nz, hz, bv, fw, free_mask, tol, sm);
}
extern "C" int fgb_extruded3_apply_jets(const fgb_extruded3_tables *xt, int32_t B, float *bvel, const float *amp, int32_t J, const int32_t *jet_face,
                                        const float *templ, int32_t nf, const float *fw, const int8_t *free_mask, float tol, fgb_stream_t s) {
    if (!xt || !bvel || !amp || !jet_face || !templ || !fw || !free_mask || B <= 0 || J <= 0 || nf <= 0)
        return set_err(FGB_E_ARG, "fgb_extruded3_apply_jets: bad argument");
    kx3_apply_jets<<<B, 512, 0, STREAM(s)>>>(xt->plane.NB, xt->nz, xt->hz, bvel, amp, J, jet_face, templ, nf, fw, free_mask, tol);
    LAUNCH_CHECK("kx3_apply_jets");
    return FGB_OK;
}
extern "C" int fgb_extruded3_balance_fluxes(const fgb_extruded3_tables *xt, int32_t B, float *bvel, const float *fw, const int8_t *free_mask,
                                            float tol, fgb_stream_t s) {
    if (!xt || !bvel || !fw || !free_mask || B <= 0) return set_err(FGB_E_ARG, "fgb_extruded3_balance_fluxes: bad argument");
    kx3_balance_fluxes<<<B, 512, 0, STREAM(s)>>>(xt->plane.NB, xt->nz, xt->hz, bvel, fw, free_mask, tol);
    LAUNCH_CHECK("kx3_balance_fluxes");
    return FGB_OK;
}
extern "C" int fgb_extruded3_update_outflow(const fgb_extruded3_tables *xt, int32_t B, const float *u, float *bvel, const float *dt, const float *fw,
                                            const int8_t *out_mask, int32_t n_out, const int32_t *out_face, const int32_t *out_cell,
                                            const float *out_adv, float tol, fgb_stream_t s) {
    if (!xt || !u || !bvel || !dt || !fw || !out_mask || !out_face || !out_cell || !out_adv || B <= 0 || n_out <= 0)
        return set_err(FGB_E_ARG, "fgb_extruded3_update_outflow: bad argument");
    kx3_update_outflow<<<B, 512, 0, STREAM(s)>>>(xt->plane.N, xt->plane.NB, xt->nz, xt->hz, u, bvel, dt, fw, out_mask, n_out, out_face, out_cell,
                                                 out_adv, tol);
    LAUNCH_CHECK("kx3_update_outflow");
    return FGB_OK;
}
extern "C" int fgb_extruded3_max_velocity(const fgb_extruded3_tables *xt, int32_t B, const float *u, const float *bvel, float *maxvel,
                                          fgb_stream_t s) {
    if (!xt || !u || !bvel || !maxvel || B <= 0) return set_err(FGB_E_ARG, "fgb_extruded3_max_velocity: bad argument");
    cudaStream_t st = STREAM(s);
    cudaError_t ce = cudaMemsetAsync(maxvel, 0, (size_t)B * sizeof(float), st);
    if (ce != cudaSuccess) return set_err(FGB_E_CUDA, "fgb_extruded3_max_velocity: memset", ce);
    X3Tab x; x.t = xt->plane; x.nz = xt->nz; x.hz = xt->hz;
    const size_t N3 = (size_t)x.t.N * x.nz;
    unsigned blocks = (unsigned)((N3 + 255) / 256);
    if (blocks > 592) blocks = 592;
    kx3_max_velocity<<<dim3(blocks, (unsigned)B), 256, 0, st>>>(x, u, bvel, maxvel);
    LAUNCH_CHECK("kx3_max_velocity");
    return FGB_OK;
}
#endif  // X3_HOST_ONLY
#endif  // __CUDACC__
