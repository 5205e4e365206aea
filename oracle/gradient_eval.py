"""TEST INFRASTRUCTURE ONLY (part of the CPU oracle; never imported by the product package fluidgym_b200).

float32 numpy restatement of the reference's cell-centred velocity gradients and of its Smagorinsky sub-grid viscosity, on the flat
neighbour-table layout of the product (``nbr [F, N]``: neighbour cell or -1 - j for prescribed face j):

* ``velocity_gradients``  getBlockDataGradient (PISO_multiblock_cuda_kernel.cu:2997-3043) as used by k_computeSpatialVelocityGradients
  (:6460-6483): per computational direction i the central difference (value_upper - value_lower) / distance with distance 2, minus 0.5
  for every side that is a prescribed (Dirichlet) face -- whose boundary value replaces the missing cell --; then the row vector of the
  computational differences times M^-1:  g_j = sum_i d_i Minv[i][j].
* ``smagorinsky_viscosity``  k_SGSviscosityIncompressibleSmagorinsky (:6913-6966): coefficient * delta * sqrt(2 S_ij S_ij) with
  S = (G + G^T) / 2 and delta = the largest squared column norm of M (the squared longest cell edge on rectilinear grids), and the
  environment's prep function (envs/tcf/tcf_env.py:441-472, envs/tcf/grid.py:75-125): times the squared van Driest factor, plus nu.

Pinned to outputs of the unmodified reference in tests/test_gradient_oracle.py: tests/golden/cyl24_velocity_gradients.npz (2-D, five
connected blocks), tests/golden/tcf32_sgs_steps.npz / tcf32_sgs_substep0.npz (3-D channel).
"""
import numpy as np

f32 = np.float32


def velocity_gradients(u, bvel, nbr, minv):
    """u [D, N], bvel [D, NB], nbr [2 D, N], minv [D, D, N] (Minv[i][j] per cell) -> G [D (component c), D (direction j), N]"""
    u, bvel = np.asarray(u, f32), np.asarray(bvel, f32)
    D, N = u.shape
    G = np.zeros((D, D, N), f32)
    for c in range(D):
        d = np.zeros((D, N), f32)
        for i in range(D):
            nl, nu = nbr[2 * i], nbr[2 * i + 1]
            lo = np.where(nl >= 0, u[c][np.maximum(nl, 0)], bvel[c][np.maximum(-1 - nl, 0)])
            hi = np.where(nu >= 0, u[c][np.maximum(nu, 0)], bvel[c][np.maximum(-1 - nu, 0)])
            dist = (f32(2.0) - np.where(nl < 0, f32(0.5), f32(0.0)) - np.where(nu < 0, f32(0.5), f32(0.0))).astype(f32)
            d[i] = ((hi - lo) / dist).astype(f32)
        for j in range(D):
            acc = np.zeros(N, f32)
            for i in range(D):
                acc = (acc + d[i] * np.asarray(minv[i][j], f32)).astype(f32)
            G[c, j] = acc
    return G


def smagorinsky_viscosity(G, h2max, coefficient, nu, damping=None):
    """G [3, 3, N] velocity gradients, h2max [N] the squared longest cell edge -> nu + C delta |S| damping"""
    S = f32(0.5) * (G + G.transpose(1, 0, 2))
    d = np.zeros(G.shape[-1], f32)
    for i in range(3):
        for j in range(i, 3):
            s = (S[i, j] * S[i, j]).astype(f32)
            d = (d + (s if i == j else f32(2.0) * s)).astype(f32)
    visc = (f32(coefficient) * np.asarray(h2max, f32) * np.sqrt(f32(2.0) * d)).astype(f32)
    if damping is not None:
        visc = (visc * np.asarray(damping, f32)).astype(f32)
    return (visc + f32(nu)).astype(f32)


def van_driest_sqr(y, u_wall, nu):
    """envs/tcf/grid.py:75-125: (1 - exp(-y+ / 25))^2 with y+ = (1 - |y|) u_wall / nu (channel centred on 0, half height 1)"""
    wd = (f32(1.0) - np.abs(np.asarray(y, f32))) * f32(u_wall) / f32(nu)
    s = f32(1.0) - np.exp(-wd * f32(1.0 / 25.0))
    return (s * s).astype(f32)
