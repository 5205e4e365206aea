"""Multi-block curvilinear domain description and its compilation into flat, batch-shared tables.

The reference keeps a mutable C++ object graph (``Domain``/``Block``/``*Boundary``,
``extensions/domain_structs.h:73-802``) that every kernel re-interprets per cell at run time
(boundary-type ``switch`` per face, connection axis permutations, corner walks).  All of that is a
function of the *geometry only* and identical for every environment of a batch, so here it is
resolved ONCE on the host into structure-of-arrays tables (neighbour indices, metric coefficients,
constant diffusion stencil, pressure-stencil weights, deferred non-orthogonal correction lists) that
the sm_100a kernels in ``csrc/`` consume for all environments of a batch at once.

Conventions (same as the reference): faces are numbered ``-x,+x,-y,+y`` = 0..3, cell index inside a
block is ``x + nx*y``, blocks are concatenated in creation order (``DS.cpp:2570-2692``), transforms
are ``[M (row major), M^-1, det]`` with column k of M = centre(face +k) - centre(face -k)
(``grid_gen.cu:298-354``); fixed-boundary transforms are slices of the face transforms
(``grid_gen.cu:398-494``, ``domain_structs.cpp:1825-1850``).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

FIXED, CONNECTED, PERIODIC = 0, 1, 2
FACE = {"-x": 0, "+x": 1, "-y": 2, "+y": 3}
f32 = np.float32


def _face(f):
    return FACE[f] if isinstance(f, str) else int(f)


@dataclass
class Boundary:
    type: int = PERIODIC
    other: int = -1
    axes: tuple = (0, 0)
    velocity: np.ndarray | None = None  # [2, n] Dirichlet values along the face (fixed only)
    scalar: np.ndarray | None = None    # [n] passive-scalar boundary value (fixed only)
    scalar_neumann: bool = False        # passive-scalar boundary condition type (False = Dirichlet)


@dataclass
class Block:
    vertex: np.ndarray  # [2, ny+1, nx+1] float32
    name: str = ""
    bounds: list = field(default_factory=lambda: [Boundary() for _ in range(4)])

    @property
    def nx(self):
        return self.vertex.shape[2] - 1

    @property
    def ny(self):
        return self.vertex.shape[1] - 1

    def size(self, axis):
        return self.nx if axis == 0 else self.ny


class DomainSpec:
    """Host-side mirror of the subset of ``PISOtorch.Domain`` the environments use to *describe* a
    domain (``CreateBlock``, ``CloseBoundary``, ``ConnectBlock``, ``MakePeriodic``; BIND.cpp:214-360)."""

    def __init__(self, viscosity: float, name: str = "domain", scalar_viscosity: float | None = None):
        self.viscosity = float(np.float32(viscosity))
        # one passive scalar channel (Domain(passiveScalarChannels=1) + setScalarViscosity, BIND.cpp:362-504)
        self.scalar_viscosity = None if scalar_viscosity is None else float(np.float32(scalar_viscosity))
        self.name = name
        self.blocks: list[Block] = []

    def create_block(self, vertex_coordinates: np.ndarray, name: str = "") -> int:
        v = np.ascontiguousarray(vertex_coordinates, dtype=np.float32)
        if v.ndim == 4:
            v = v[0]
        assert v.ndim == 3 and v.shape[0] == 2
        self.blocks.append(Block(v, name))
        return len(self.blocks) - 1

    def close_boundary(self, block: int, face, velocity=None, scalar=None, scalar_neumann=False):
        f = _face(face)
        b = self.blocks[block]
        # CloseConnectedBoudary (domain_structs.cpp:1789-1823): closing one side of a periodic pair (the default
        # state of every face) or of a block connection also closes its partner face with a zero Dirichlet value
        prev = b.bounds[f]
        if prev.type == PERIODIC:
            n_ = b.size(1 - (f >> 1))
            b.bounds[f ^ 1] = Boundary(FIXED, velocity=np.zeros((2, n_), dtype=np.float32),
                                       scalar=np.zeros(n_, dtype=np.float32), scalar_neumann=False)
        elif prev.type == CONNECTED:
            ob = self.blocks[prev.other]
            of = prev.axes[0]
            n_ = ob.size(1 - (of >> 1))
            ob.bounds[of] = Boundary(FIXED, velocity=np.zeros((2, n_), dtype=np.float32), scalar=np.zeros(n_, dtype=np.float32),
                                     scalar_neumann=False)
        n = b.size(1 - (f >> 1))
        vel = np.zeros((2, n), dtype=np.float32)
        if velocity is not None:
            velocity = np.asarray(velocity, dtype=np.float32).reshape(2, -1)
            vel[:] = velocity  # broadcasts a static [2,1] value
        sc = np.zeros(n, dtype=np.float32)
        if scalar is not None:
            sc[:] = np.asarray(scalar, dtype=np.float32).reshape(-1)
        b.bounds[f] = Boundary(FIXED, velocity=vel, scalar=sc, scalar_neumann=bool(scalar_neumann))

    def connect(self, block1: int, face1, block2: int, face2, axis1):
        """``ConnectBlocks`` (domain_structs.cpp:1080-1113), 2-D case."""
        f1, f2, a1 = _face(face1), _face(face2), _face(axis1)
        self.blocks[block1].bounds[f1] = Boundary(CONNECTED, block2, (f2, a1))
        back = ((((f1 >> 1) + 1) % 2) << 1) | (a1 & 1)
        self.blocks[block2].bounds[f2] = Boundary(CONNECTED, block1, (f1, back))

    def make_periodic(self, block: int, axis: int):
        self.blocks[block].bounds[2 * axis] = Boundary(PERIODIC)
        self.blocks[block].bounds[2 * axis + 1] = Boundary(PERIODIC)

    # ------------------------------------------------------------------
    def prepare(self) -> "CompiledDomain":
        return CompiledDomain(self)


# ----------------------------------------------------------------------
def cell_transforms(vertex: np.ndarray) -> np.ndarray:
    """[ny, nx, 9] = M(4), Minv(4), det in float32 (grid_gen.cu:298-354)."""
    v = vertex.astype(f32)
    half = f32(0.5)
    fxm = (v[:, :-1, :-1] + v[:, 1:, :-1]) * half  # -x face centre
    fxp = (v[:, :-1, 1:] + v[:, 1:, 1:]) * half
    fym = (v[:, :-1, :-1] + v[:, :-1, 1:]) * half
    fyp = (v[:, 1:, :-1] + v[:, 1:, 1:]) * half
    m00 = fxp[0] - fxm[0]
    m10 = fxp[1] - fxm[1]
    m01 = fyp[0] - fym[0]
    m11 = fyp[1] - fym[1]
    return _pack_transform(m00, m01, m10, m11)


def _pack_transform(m00, m01, m10, m11):
    det = (m00 * m11 - m10 * m01).astype(f32)
    r = (f32(1.0) / det).astype(f32)
    T = np.stack([m00, m01, m10, m11, m11 * r, -m01 * r, -m10 * r, m00 * r, det], axis=-1)
    return np.ascontiguousarray(T, dtype=f32)


def boundary_transforms(vertex: np.ndarray, face: int) -> np.ndarray:
    """[n, 9] face transforms on a block boundary (grid_gen.cu:398-494 restricted to the slice that
    ``Block::GetFaceTransformBoundarySlice`` takes, domain_structs.cpp:1825-1850)."""
    v = vertex.astype(f32)
    axis, upper = face >> 1, face & 1
    if axis == 0:  # x-normal faces on vertex column 0 / nx ; tangential = y
        col = v.shape[2] - 1 if upper else 0
        inner = col - 1 if upper else col + 1
        sgn = f32(1.0) if upper else f32(-1.0)
        # normal column: (sum of the two face vertices one column inward) with sign, one-sided, weight 1/2
        own = v[:, :-1, col] + v[:, 1:, col]
        nb = v[:, :-1, inner] + v[:, 1:, inner]
        if upper:
            n_vec = (own - nb)
        else:
            n_vec = (nb - own)
        n_vec = n_vec * f32(0.5)
        t_vec = v[:, 1:, col] - v[:, :-1, col]
        m00, m10 = n_vec[0], n_vec[1]
        m01, m11 = t_vec[0], t_vec[1]
        del sgn
    else:
        row = v.shape[1] - 1 if upper else 0
        inner = row - 1 if upper else row + 1
        own = v[:, row, :-1] + v[:, row, 1:]
        nb = v[:, inner, :-1] + v[:, inner, 1:]
        n_vec = (own - nb) if upper else (nb - own)
        n_vec = n_vec * f32(0.5)
        t_vec = v[:, row, 1:] - v[:, row, :-1]
        m01, m11 = n_vec[0], n_vec[1]
        m00, m10 = t_vec[0], t_vec[1]
    return _pack_transform(m00, m01, m10, m11)


# ----------------------------------------------------------------------
class CompiledDomain:
    """Flat tables for one geometry (shared by every environment of a batch)."""

    def __init__(self, spec: DomainSpec, transforms=None, btransforms=None, conn_corner_offset: int = 1):
        self.spec = spec
        self.visc = f32(spec.viscosity)
        self.conn_corner_offset = conn_corner_offset
        blocks = spec.blocks
        self.nb = len(blocks)
        self.sizes = np.array([[b.nx, b.ny] for b in blocks], dtype=np.int32)
        self.offsets = np.concatenate([[0], np.cumsum(self.sizes[:, 0] * self.sizes[:, 1])]).astype(np.int64)
        self.N = int(self.offsets[-1])
        self.btype = np.array([[bd.type for bd in b.bounds] for b in blocks], dtype=np.int32)
        self.bconn = np.array([[[bd.other, bd.axes[0], bd.axes[1]] for bd in b.bounds] for b in blocks], dtype=np.int32)
        # boundary-face numbering: blocks in order, faces in order, tangential index
        self.boff = -np.ones((self.nb, 4), dtype=np.int64)
        nbf = 0
        for bi, b in enumerate(blocks):
            for f in range(4):
                if b.bounds[f].type == FIXED:
                    self.boff[bi, f] = nbf
                    nbf += b.size(1 - (f >> 1))
        self.NB = nbf
        # transforms
        if transforms is None:
            transforms = [cell_transforms(b.vertex) for b in blocks]
        self.T = np.concatenate([t.reshape(-1, 9) for t in transforms]).astype(f32)
        bT = np.zeros((max(self.NB, 1), 9), dtype=f32)
        bvel = np.zeros((2, max(self.NB, 1)), dtype=f32)
        self.b_cell = np.zeros(max(self.NB, 1), dtype=np.int32)
        self.b_face = np.zeros(max(self.NB, 1), dtype=np.int32)
        self.sb_val0 = np.zeros(max(self.NB, 1), dtype=f32)
        self.sb_neumann = np.zeros(max(self.NB, 1), dtype=np.int8)
        self.scalar_visc = None if spec.scalar_viscosity is None else f32(spec.scalar_viscosity)
        for bi, b in enumerate(blocks):
            for f in range(4):
                if b.bounds[f].type != FIXED:
                    continue
                n = b.size(1 - (f >> 1))
                o = self.boff[bi, f]
                if btransforms is not None and (bi, f) in btransforms:
                    bT[o:o + n] = btransforms[(bi, f)].reshape(-1, 9)
                else:
                    bT[o:o + n] = boundary_transforms(b.vertex, f)
                bvel[:, o:o + n] = b.bounds[f].velocity
                if b.bounds[f].scalar is not None:
                    self.sb_val0[o:o + n] = b.bounds[f].scalar
                    self.sb_neumann[o:o + n] = 1 if b.bounds[f].scalar_neumann else 0
                for k in range(n):
                    pos = [0, 0]
                    pos[1 - (f >> 1)] = k
                    pos[f >> 1] = b.size(f >> 1) - 1 if (f & 1) else 0
                    self.b_cell[o + k] = self.gidx(bi, pos)
                    self.b_face[o + k] = f
        self.bT = bT
        self.bvel0 = bvel
        self._build_tables()

    # ---- index helpers (mirror K.cu:152-199, 329-375) -------------------------------------------
    def gidx(self, bi, pos):
        return int(self.offsets[bi] + pos[0] + self.sizes[bi, 0] * pos[1])

    def at_bound(self, bi, pos, f):
        ax = f >> 1
        return pos[ax] == self.sizes[bi, ax] - 1 if (f & 1) else pos[ax] == 0

    def bface(self, bi, f, pos):
        return int(self.boff[bi, f] + pos[1 - (f >> 1)])

    def connected_pos(self, bi, f, pos, border_offset):
        ob, a0, a1 = self.bconn[bi, f]
        out = [0, 0]
        ca = a0 >> 1
        out[ca] = self.sizes[ob, ca] - 1 - border_offset if (a0 & 1) else border_offset
        axis = ((f >> 1) + 1) % 2
        ca = a1 >> 1
        out[ca] = self.sizes[ob, ca] - 1 - pos[axis] if (a1 & 1) else pos[axis]
        return int(ob), out

    def connected_dir(self, bi, f, d):
        rel = ((d >> 1) - (f >> 1)) % 2
        return int(self.bconn[bi, f, 1 + rel]) ^ (d & 1)

    def neighbor(self, bi, pos, f, border_offset=0):
        """-> (block, pos) of the cell across face f, or None for a prescribed boundary."""
        ax = f >> 1
        if self.at_bound(bi, pos, f):
            t = self.btype[bi, f]
            if t == FIXED:
                return None
            if t == CONNECTED:
                return self.connected_pos(bi, f, pos, border_offset)
            p = list(pos)
            p[ax] = 0 if (f & 1) else self.sizes[bi, ax] - 1
            return bi, p
        p = list(pos)
        p[ax] += (f & 1) * 2 - 1
        return bi, p

    def corner(self, bi, pos, dir1, dir2):
        """getCornerValue walk (K.cu:2757-2874) with includeDepth0/1 = False, maxDepth = 2.
        Returns (num_cells, [cells at depth 2]) or (0, [boundary faces])."""
        num = 1
        cells = []
        cy = [[dir1, dir2, bi, list(pos)], [dir2, dir1, bi, list(pos)]]
        for depth in (1, 2):
            for k in range(2):
                d1, d2, cb, p = cy[k]
                ax = d1 >> 1
                fs = (d1 & 1) * 2 - 1
                if self.at_bound(cb, p, d1):
                    t = self.btype[cb, d1]
                    if t == FIXED:
                        bf = [self.bface(cb, d1, p)]
                        if not self.at_bound(cb, p, d2):
                            q = list(p)
                            q[d2 >> 1] += (d2 & 1) * 2 - 1
                            bf.append(self.bface(cb, d1, q))
                        return 0, bf
                    if t == CONNECTED:
                        nb, npos = self.connected_pos(cb, d1, p, self.conn_corner_offset)
                        nd1 = self.connected_dir(cb, d1, d2)
                        nd2 = self.connected_dir(cb, d1, d1) ^ 1
                        cy[k] = [nd1, nd2, nb, npos]
                    else:
                        p = list(p)
                        p[ax] = 0 if (d1 & 1) else self.sizes[cb, ax] - 1
                        cy[k] = [d2, d1 ^ 1, cb, p]
                else:
                    p = list(p)
                    p[ax] += fs
                    cy[k] = [d2, d1 ^ 1, cb, p]
                o = cy[k ^ 1]
                if cy[k][2] == o[2] and cy[k][3] == o[3]:
                    return num, cells
                if depth > 1:
                    cells.append(self.gidx(cy[k][2], cy[k][3]))
                num += 1
        return num, cells

    def diag_neighbor(self, bi, pos, dir1, dir2):
        """getBlockDataNeighborDiagonal (K.cu:2629-2677): global cell or -1-bface."""
        e1 = self.btype[bi, dir1] == FIXED
        dirs = (dir2, dir1) if e1 else (dir1, dir2)
        cb, p = bi, list(pos)
        for f in dirs:
            ax = f >> 1
            if self.at_bound(cb, p, f):
                t = self.btype[cb, f]
                if t == FIXED:
                    return -1 - self.bface(cb, f, p)
                if t == CONNECTED:
                    cb, p = self.connected_pos(cb, f, p, self.conn_corner_offset)
                else:
                    p[ax] = 0 if (f & 1) else self.sizes[cb, ax] - 1
            else:
                p[ax] += (f & 1) * 2 - 1
        return self.gidx(cb, p)

    # ---------------------------------------------------------------------------------------------
    def _build_tables(self):
        N, NB = self.N, self.NB
        T = self.T
        det = T[:, 8]
        mi = T[:, 4:8]
        self.det = det.copy()
        self.minv = np.ascontiguousarray(mi.T)  # [4, N]
        a00 = (det * (mi[:, 0] * mi[:, 0] + mi[:, 1] * mi[:, 1])).astype(f32)
        a11 = (det * (mi[:, 2] * mi[:, 2] + mi[:, 3] * mi[:, 3])).astype(f32)
        a01 = (det * (mi[:, 2] * mi[:, 0] + mi[:, 3] * mi[:, 1])).astype(f32)
        alpha = np.stack([a00, a11])  # [2, N]
        self.alpha = alpha
        self.alpha01 = a01
        bT = self.bT
        bdet, bmi = bT[:, 8], bT[:, 4:8]
        b00 = (bdet * (bmi[:, 0] * bmi[:, 0] + bmi[:, 1] * bmi[:, 1])).astype(f32)
        b11 = (bdet * (bmi[:, 2] * bmi[:, 2] + bmi[:, 3] * bmi[:, 3])).astype(f32)
        b01 = (bdet * (bmi[:, 2] * bmi[:, 0] + bmi[:, 3] * bmi[:, 1])).astype(f32)
        self.b_det = bdet.copy()
        self.b_minv = np.ascontiguousarray(bmi.T)
        ax_b = self.b_face >> 1
        self.b_alpha = np.where(ax_b == 0, b00, b11).astype(f32)  # alpha_b^{dd} of the face normal axis

        nbr = np.full((4, N), -1, dtype=np.int32)
        fl_comp = np.zeros((4, N), dtype=np.int8)   # bit0: neighbour component, bit1: negate
        nalpha = np.zeros((4, N), dtype=f32)        # alpha_N^{d'd'} (mapped component)
        Cd = np.zeros((5, N), dtype=f32)            # constant (diffusive) part of C before /det
        Wp = np.zeros((5, 5, N), dtype=f32)         # P_e = sum_j Wp[e][j] * rA_j
        no_entries = [[] for _ in range(N)]         # (cell j, face, gP, gN)
        nob_entries = [[] for _ in range(N)]        # (bface j, weight) velocity only (weight excludes nu)
        visc = self.visc
        half = f32(0.5)

        for bi in range(self.nb):
            nx, ny = self.sizes[bi]
            for y in range(ny):
                for x in range(nx):
                    pos = [x, y]
                    g = self.gidx(bi, pos)
                    # --- neighbour resolution
                    nb_g = [-1] * 4
                    ncell = [None] * 4
                    for f in range(4):
                        r = self.neighbor(bi, pos, f)
                        if r is None:
                            nbr[f, g] = -1 - self.bface(bi, f, pos)
                            continue
                        gn = self.gidx(*r)
                        nb_g[f] = gn
                        ncell[f] = r
                        nbr[f, g] = gn
                        dim = f >> 1
                        ch = dim
                        neg = 0
                        if self.at_bound(bi, pos, f) and self.btype[bi, f] == CONNECTED:
                            a0 = self.bconn[bi, f, 1]
                            ch = a0 >> 1
                            neg = 1 if (a0 & 1) == (f & 1) else 0
                        fl_comp[f, g] = ch | (neg << 1)
                        nalpha[f, g] = alpha[ch, gn]
                    # --- interpolated non-orthogonal face coefficients, matrix flavour (K.cu:1926-2001):
                    # neighbour alpha01 WITHOUT axis mapping; zero on faces whose block boundary is FIXED
                    # unless the cell is strictly interior along that axis.
                    aP01 = a01[g]
                    mat_face = [None] * 4  # (aP01, aN01) or None
                    for f in range(4):
                        ax = f >> 1
                        interior = 0 < pos[ax] < self.sizes[bi, ax] - 1
                        if (interior or self.btype[bi, f] != FIXED) and nb_g[f] >= 0:
                            mat_face[f] = (aP01, a01[nb_g[f]])
                    # --- corners
                    corner = {}
                    for f in range(4):
                        if nb_g[f] < 0:
                            continue
                        tax = ((f >> 1) + 1) % 2
                        for tu in range(2):
                            tf = (tax << 1) | tu
                            corner[(f, tf)] = self.corner(bi, pos, f, tf)
                    # --- constant diffusion stencil (K.cu:3692-3848) and pressure weights (K.cu:4842-4952)
                    diag = f32(0.0)
                    off = [f32(0.0)] * 4
                    for f in range(4):
                        dim = f >> 1
                        fs = f32((f & 1) * 2 - 1)
                        if nb_g[f] < 0:
                            diag = f32(diag + f32(2.0) * visc * alpha[dim, g])
                            continue
                        vc = f32((alpha[dim, g] * visc + nalpha[f, g] * visc) * half)
                        diag = f32(diag + vc)
                        off[f] = f32(off[f] - vc)
                        # pressure orthogonal part: 0.5*(alphaP*raP + alphaN*raN)
                        Wp[0, 0, g] -= half * alpha[dim, g]
                        Wp[0, f + 1, g] -= half * nalpha[f, g]
                        Wp[f + 1, 0, g] += half * alpha[dim, g]
                        Wp[f + 1, f + 1, g] += half * nalpha[f, g]
                        if mat_face[f] is None:
                            continue
                        aP, aN = mat_face[f]
                        alpha_v = f32((aP * visc + aN * visc) * half)
                        if alpha_v == 0 and aP == 0 and aN == 0:
                            continue
                        tax = (dim + 1) % 2
                        for tu in range(2):
                            tf = (tax << 1) | tu
                            tfs = f32(tu * 2 - 1)
                            num, _ = corner[(f, tf)]
                            if num < 1:
                                # velocity: Dirichlet corner -> RHS only. pressure: one-sided (K.cu:4913-4929)
                                cP = fs * tfs * half * aP * f32(0.25)
                                cN = fs * tfs * half * aN * f32(0.25)
                                for e, s in ((0, 3.0), (f + 1, 3.0), ((tf ^ 1) + 1, -1.0)):
                                    Wp[e, 0, g] += f32(s) * cP
                                    Wp[e, f + 1, g] += f32(s) * cN
                            else:
                                inv = f32(1.0) / f32(num)
                                if alpha_v != 0:
                                    c = f32(fs * tfs * alpha_v * inv)
                                    diag = f32(diag - c)
                                    off[f] = f32(off[f] - c)
                                    off[tf & 3] = f32(off[tf] - c)
                                cP = fs * tfs * half * aP * inv
                                cN = fs * tfs * half * aN * inv
                                for e in (0, f + 1, tf + 1):
                                    Wp[e, 0, g] += cP
                                    Wp[e, f + 1, g] += cN
                    Cd[0, g] = diag
                    for f in range(4):
                        Cd[f + 1, g] = off[f] if nb_g[f] >= 0 else 0.0
                        if nb_g[f] < 0:
                            Wp[f + 1, :, g] = 0.0
                    # --- deferred non-orthogonal terms (K.cu:3048-3202)
                    for f in range(4):
                        ax = f >> 1
                        fs = f32((f & 1) * 2 - 1)
                        tax = (ax + 1) % 2
                        if nb_g[f] < 0:
                            # prescribed face: tangential gradient of the boundary data (velocity only)
                            j0 = self.bface(bi, f, pos)
                            balpha = b01[j0]
                            tl = pos[tax] == 0
                            tuu = pos[tax] == self.sizes[bi, tax] - 1
                            lo, up = list(pos), list(pos)
                            df = f32(0.5)
                            if not tl:
                                lo[tax] -= 1
                            if not tuu:
                                up[tax] += 1
                            if tl or tuu:
                                df = f32(1.0)
                            w = f32(-fs * balpha * df)
                            nob_entries[g].append((self.bface(bi, f, up), w))
                            nob_entries[g].append((self.bface(bi, f, lo), f32(-w)))
                            continue
                        # neighbour alpha01 WITH axis mapping (K.cu:1430-1466)
                        gn = nb_g[f]
                        m_ax, m_t = ax, tax
                        if self.at_bound(bi, pos, f) and self.btype[bi, f] == CONNECTED:
                            m_ax = self.connected_dir(bi, f, ax << 1) >> 1
                            m_t = self.connected_dir(bi, f, tax << 1) >> 1
                        aN = self._alpha_pair(gn, m_t, m_ax)
                        aP = a01[g]
                        for tu in range(2):
                            tf = (tax << 1) | tu
                            tfs = f32(tu * 2 - 1)
                            num, items = corner[(f, tf)]
                            if num == 0:
                                # velocity (Dirichlet): boundary value(s); pressure: diagonal one-sided
                                wb = f32(1.0) if len(items) == 1 else f32(0.5)
                                for j in items:
                                    nob_entries[g].append((j, f32(-fs * tfs * wb), f, aP, aN))
                                dn = self.diag_neighbor(bi, pos, f, tf ^ 1)
                                if dn >= 0:
                                    s = f32(-fs * (-tfs) * f32(0.25))
                                    no_entries[g].append((dn, f, f32(s * half * aP), f32(s * half * aN), 1))
                            else:
                                inv = f32(1.0) / f32(num)
                                for j in items:
                                    s = f32(-fs * tfs * inv)
                                    no_entries[g].append((j, f, f32(s * half * aP), f32(s * half * aN), 0))
        # constant part of the passive-scalar transport matrix (K.cu:3692-3750, 3816-3848 with
        # forPassiveScalar): orthogonal diffusion with the scalar diffusivity, Dirichlet faces add 2*kappa*alpha.
        # (The deferred non-orthogonal scalar terms are not tabulated: the only scalar environments, RBC, run
        # with non_orthogonal=False on an orthogonal grid, rbc_env_base.py:306-329.)
        kap = self.scalar_visc if self.scalar_visc is not None else f32(0.0)
        Cd_s = np.zeros((5, N), dtype=f32)
        for f in range(4):
            dim = f >> 1
            inner = nbr[f] >= 0
            vc = ((alpha[dim] * kap + nalpha[f] * kap) * half).astype(f32)
            j = np.where(inner, 0, -1 - nbr[f])
            dirichlet = (self.sb_neumann[j] == 0)
            Cd_s[0] += np.where(inner, vc, np.where(dirichlet, f32(2.0) * kap * alpha[dim], f32(0.0))).astype(f32)
            Cd_s[f + 1] = np.where(inner, -vc, f32(0.0))
        self.Cd_s = Cd_s
        # reverse faces: rev[f][g] = face f' of the neighbour n = nbr[f][g] with nbr[f'][n] == g
        rev = np.full((4, N), -1, dtype=np.int8)
        cells = np.arange(N)
        for f in range(4):
            inner = nbr[f] >= 0
            nn = np.where(inner, nbr[f], 0)
            for f2 in range(4):
                hit = inner & (nbr[f2][nn] == cells) & (rev[f] < 0)
                rev[f][hit] = f2
            assert (rev[f][inner] >= 0).all(), "neighbour relation is not symmetric"
        self.rev = rev
        self.nbr = nbr
        self.fl_comp = fl_comp
        self.nalpha = nalpha
        self.Cd = Cd
        self.Wp = Wp
        # pack ELL lists.  no_*: kind 0 -> used by velocity and pressure; kind 1 -> pressure only
        K = max(1, max(len(e) for e in no_entries))
        self.no_idx = np.zeros((K, N), dtype=np.int32)
        self.no_face = np.zeros((K, N), dtype=np.int8)
        self.no_gP = np.zeros((K, N), dtype=f32)
        self.no_gN = np.zeros((K, N), dtype=f32)
        self.no_wv = np.zeros((K, N), dtype=f32)   # velocity weight (nu folded in), 0 for pressure-only entries
        for g, lst in enumerate(no_entries):
            for k, (j, f, gP, gN, kind) in enumerate(lst):
                self.no_idx[k, g] = j
                self.no_face[k, g] = f
                self.no_gP[k, g] = gP
                self.no_gN[k, g] = gN
                if kind == 0:
                    self.no_wv[k, g] = f32(f32(gP * visc) + f32(gN * visc))
        Kb = max(1, max(len(e) for e in nob_entries))
        self.nob_idx = np.zeros((Kb, N), dtype=np.int32)
        self.nob_w = np.zeros((Kb, N), dtype=f32)
        for g, lst in enumerate(nob_entries):
            for k, e in enumerate(lst):
                if len(e) == 2:
                    j, w = e
                    self.nob_idx[k, g] = j
                    self.nob_w[k, g] = f32(w * visc)
                else:
                    j, s, f, aP, aN = e
                    fa = f32((aP * visc + aN * visc) * half)
                    self.nob_idx[k, g] = j
                    self.nob_w[k, g] = f32(s * fa)
        self.K_no, self.K_nob = K, Kb

    def _alpha_pair(self, g, c1, c2):
        mi, det = self.T[g, 4:8], self.T[g, 8]
        r1, r2 = mi[2 * c1:2 * c1 + 2], mi[2 * c2:2 * c2 + 2]
        return f32(det * f32(f32(r1[0] * r2[0]) + f32(r1[1] * r2[1])))
