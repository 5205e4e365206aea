"""Static plan of the register-blocked on-chip pressure CG (``cg_impl`` 11, ``k_cg_strip`` in csrc/piso_b200.cu).

The reference solves the pressure system with cuSPARSE SpMV + cuBLAS dot products on CSR (CG.cu:225-446).  Here one
thread-block cluster owns one environment and every CTA keeps its part of the system on chip.  This module lays
the cells of a CTA out as padded 2-D arrays ("segments") so that

* the four stencil neighbours of slot ``s`` are ``s - 1, s + 1, s - S, s + S`` (S = row stride of the segment): no
  per-cell neighbour addresses, the kernel needs one base address and one stride per thread;
* a thread owns a column run of CPT consecutive rows: the north / south neighbours of its cells are its own
  registers, only the east / west neighbours and the two run ends are shared-memory loads;
* S is padded so that (CPT - 1) S = 0 mod 32: the slot of lane l+1 then always lies one bank after the slot of lane l,
  also where a warp runs from one band of rows into the next (CPT S = S mod 32): every access is conflict free
  (CPT = 13: S a multiple of 8, CPT = 9: of 4, CPT = 17: of 2);
* whatever is not array-adjacent (block connections, periodic wrap, cuts between CTAs) is a GHOST slot.
  A REMOTE ghost (owner in another CTA) is a full replica of the remote cell: it runs the same x / p updates with
  the same scalars as its owner, so it stays bit-identical to it; the only value exchanged per iteration is the
  residual r of the mirrored cell, pushed by its owner together with the partial sums of <r, r>.
  A LOCAL ghost (owner in the same CTA) is a mirror: right after storing its rows the owning thread copies the
  mirrored ones into the ghost slots (shared memory to shared memory, before the CTA barrier that publishes the vector).

Layout directions are W, E (along a row, slot -/+ 1) and S, N (across rows, slot -/+ S); ``fmap`` maps them to the
mesh faces (0: -x, 1: +x, 2: -y, 3: +y) of the block the segment was cut from.
"""
from __future__ import annotations

import numpy as np

DEAD = -1
import math
import os

CS_CHOICES = (1, 2, 4, 8, 16)
SHAPES = ((480, 17), (256, 17), (896, 9), (640, 13), (320, 13))   # (threads per CTA, rows per thread) the kernel is instantiated for (1 / 2 co-resident
#                                   CTAs per SM); the first that fits is the default.  Threads per CTA: a multiple of 32.


SHAPES2 = ((480, 9),)       # shapes of k_cg_strip2 (cg_impl 12: two environments per cluster, each with these rows per thread)


def pair_shape():
    """(T, CPT) of the two-environment kernel; FGB_STRIP2_SHAPE="T,CPT" selects another instantiated shape."""
    v = os.environ.get("FGB_STRIP2_SHAPE")
    if v:
        T, cpt = (int(x) for x in v.split(","))
        assert (T, cpt) in SHAPES2, f"FGB_STRIP2_SHAPE must be one of {SHAPES2}"
        return T, cpt
    return SHAPES2[0]


def default_shape():
    """(T, CPT) of the plan; FGB_STRIP_SHAPE="T,CPT" selects another instantiated shape (A/B runs)."""
    v = os.environ.get("FGB_STRIP_SHAPE")
    if v:
        T, cpt = (int(x) for x in v.split(","))
        assert (T, cpt) in SHAPES, f"FGB_STRIP_SHAPE must be one of {SHAPES}"
        return T, cpt
    return SHAPES[0]


THREADS, CPT = SHAPES[0]
TH_FIELDS = 8     # per-thread record: base slot, stride, fmap, remote exports (offset | count << 16), local exports (same),
#                   index of the first remote ghost in the receive buffer, remote ghost rows (bit k), local ghost rows (bit k)


def ghost_code(g: int) -> int:
    return -2 - int(g)


class StripPlan:
    """cs, T, cpt, slots (per CTA, incl. padding); thread[cs][T][8]; cell[cs][T][cpt] (cell id, -1 dead, -2-g ghost);
    rexp[cs][remax][2] = (k | dst rank << 8, dst index in the receive buffer); lexp[cs][lemax][2] = (src slot, dst slot);
    cnt[cs][4] = (remote ghosts, remote exports, local exports, 0)."""

    def __init__(self, cs, T, cpt, slots, thread, cell, rexp, lexp, cnt, segments):
        self.cs, self.T, self.cpt, self.slots = cs, T, cpt, slots
        self.thread, self.cell, self.rexp, self.lexp, self.cnt = thread, cell, rexp, lexp, cnt
        self.remax, self.lemax = rexp.shape[1], lexp.shape[1]
        self.gmax = int(cnt[:, 0].max()) if cs else 0
        self.segments = segments

    def stats(self):
        return dict(cs=self.cs, T=self.T, cpt=self.cpt, slots=self.slots, real=int((self.cell >= 0).sum()),
                    ghost=int((self.cell <= -2).sum()), thread_slots=self.cs * self.T * self.cpt,
                    remote_ghosts=self.cnt[:, 0].tolist(), remote_exports=self.cnt[:, 1].tolist(), local_exports=self.cnt[:, 2].tolist(),
                    active_threads=[int((self.thread[r, :, 1] > 0).sum()) for r in range(self.cs)])


def _block_axes(sizes, nbr, offsets, bi):
    """Column axis of block bi: the axis with fewer connected end faces (no ghost columns when both ends are
    boundaries), ties -> the shorter one."""
    nx, ny = int(sizes[bi][0]), int(sizes[bi][1])
    off = int(offsets[bi])
    cells = off + np.arange(nx * ny).reshape(ny, nx)
    conn = []
    for f, edge in ((0, cells[:, 0]), (1, cells[:, -1]), (2, cells[0, :]), (3, cells[-1, :])):
        conn.append(bool((nbr[f][edge] >= 0).any()))
    cx, cy = conn[0] + conn[1], conn[2] + conn[3]
    if cx != cy:
        axis = 0 if cx < cy else 1
    else:
        axis = 0 if nx <= ny else 1
    return axis, (conn[0] or conn[1]) if axis == 0 else (conn[2] or conn[3])


def _try_plan(sizes, offsets, nbr, N, cs, T, cpt):
    nb = len(sizes)
    segs = [[] for _ in range(cs)]      # per CTA: (block, axis, r0, n_rows, W, S, ghost_cols, bands)
    used = [0] * cs
    cta = 0
    for bi in range(nb):
        axis, gcols = _block_axes(sizes, nbr, offsets, bi)
        nx, ny = int(sizes[bi][0]), int(sizes[bi][1])
        W, R = (nx, ny) if axis == 0 else (ny, nx)
        if W + 2 > T // 2:                       # too wide for a row of threads: take the other axis
            axis, W, R = 1 - axis, R, W
            gcols = True
        S = W + (2 if gcols else 0)
        S += (-S) % (32 // math.gcd(32, cpt - 1))   # conflict-free band transitions (module docstring)
        r0 = 0
        while r0 < R:
            if cta >= cs:
                return None
            bands_fit = (T - used[cta]) // S
            n = min(R - r0, bands_fit * cpt - 2)
            if bands_fit == 0 or n <= 0 or (n < R - r0 and n < cpt):   # do not leave slivers
                cta += 1
                continue
            bands = -(-(n + 2) // cpt)
            segs[cta].append((bi, axis, r0, n, W, S, gcols, bands))
            used[cta] += bands * S
            r0 += n
    return segs


def build_strip_plan(sizes, offsets, nbr, N, cs=None, T=None, cpt=None):
    """Returns a StripPlan or None when the domain does not fit the available cluster sizes."""
    if T is None or cpt is None:
        T, cpt = default_shape()
    sizes = np.asarray(sizes)
    nbr = np.asarray(nbr)
    pick = None
    for c in ((cs,) if cs else CS_CHOICES):
        if N > c * T * cpt:
            continue
        segs = _try_plan(sizes, offsets, nbr, N, c, T, cpt)
        if segs is not None:
            pick = (c, segs)
            break
    if pick is None:
        return None
    cs, segs = pick
    smax = max(s[5] for per in segs for s in per)
    # slot layout per CTA: [smax zero pad][segment 0: bands*cpt rows x S] ... [smax zero pad]
    slot_of_cell = np.full(N, -1, dtype=np.int64)
    thread = np.zeros((cs, T, TH_FIELDS), dtype=np.int32)
    thread[:, :, 0] = 1                                  # idle threads: base 1, stride 0 (inside the leading zero pad)
    cell = np.full((cs, T, cpt), DEAD, dtype=np.int32)
    nslots = []
    seg_geo = []
    for r in range(cs):
        base = smax
        t = 0
        for (bi, axis, r0, n, W, S, gcols, bands) in segs[r]:
            nx = int(sizes[bi][0])
            off = int(offsets[bi])
            c0 = 1 if gcols else 0
            fmap = (0 | (1 << 2) | (2 << 4) | (3 << 6)) if axis == 0 else (2 | (3 << 2) | (0 << 4) | (1 << 6))   # W, E, S, N
            seg_geo.append((r, base, bi, axis, r0, n, W, S, c0, bands, t))
            for band in range(bands):
                for c in range(S):
                    thread[r, t, 0:3] = (base + band * cpt * S + c, S, fmap)
                    for k in range(cpt):
                        row = band * cpt + k - 1          # array row 0 is the ghost row below the segment
                        col = c - c0
                        if 0 <= row < n and 0 <= col < W:
                            j, i = r0 + row, col
                            g = off + (j * nx + i if axis == 0 else i * nx + j)
                            cell[r, t, k] = g
                            slot_of_cell[g] = base + (band * cpt + k) * S + c
                    t += 1
            base += bands * cpt * S
        nslots.append(base + smax)
    assert (slot_of_cell >= 0).all(), "strip plan does not cover every cell"
    slots = int(max(nslots))
    slots += (-slots) % 4
    # ghost slots: array positions adjacent to a real cell whose mesh neighbour is not array-adjacent
    dirs = ((-1, 0), (1, 0), (0, -1), (0, 1))            # W, E, S, N as (dcol, drow)
    for (r, base, bi, axis, r0, n, W, S, c0, bands, t0) in seg_geo:
        fm = (0, 1, 2, 3) if axis == 0 else (2, 3, 0, 1)
        rows = bands * cpt
        view = cell[r, t0:t0 + bands * S].reshape(bands, S, cpt).transpose(0, 2, 1).reshape(rows, S)   # [array row][col]
        for arow in range(rows):
            for c in range(S):
                g = view[arow, c]
                if g < 0:
                    continue
                for d, (dc, dr) in enumerate(dirs):
                    nbg = int(nbr[fm[d]][g])
                    if nbg < 0:
                        continue
                    ar, ac = arow + dr, c + dc
                    assert 0 <= ar < rows and 0 <= ac < S, "mesh neighbour falls outside the padded segment"
                    cur = view[ar, ac]
                    if cur >= 0:
                        assert cur == nbg, "array neighbour is not the mesh neighbour"
                    elif cur == DEAD:
                        view[ar, ac] = ghost_code(nbg)
                    else:
                        assert cur == ghost_code(nbg), "two different cells need the same ghost slot"
        cell[r, t0:t0 + bands * S] = view.reshape(bands, cpt, S).transpose(0, 2, 1).reshape(bands * S, cpt)
    # owners of the mirrored cells; ghost rows: remote -> index in the receive buffer, local -> mirror slot
    src_thread = {}
    for r in range(cs):
        tt, kk = np.nonzero(cell[r] >= 0)
        for t, k in zip(tt, kk):
            src_thread[int(cell[r, t, k])] = (r, int(t), int(k))
    rexports = [dict() for _ in range(cs)]               # src rank -> {src thread: [(k, dst rank, dst index)]}
    lexports = [dict() for _ in range(cs)]               # src rank -> {src thread: [(k, dst slot)]}
    n_rghost = np.zeros(cs, dtype=np.int64)
    for r in range(cs):
        for t in range(T):
            rm = lm = 0
            first = int(n_rghost[r])
            for k in range(cpt):
                c = int(cell[r, t, k])
                if c > -2:
                    continue
                sr, st, sk = src_thread[-2 - c]
                if sr == r:
                    lm |= 1 << k
                    lexports[r].setdefault(st, []).append((int(thread[r, st, 0]) + sk * int(thread[r, st, 1]),
                                                           int(thread[r, t, 0]) + k * int(thread[r, t, 1])))
                else:
                    rm |= 1 << k
                    rexports[sr].setdefault(st, []).append((sk, r, int(n_rghost[r])))
                    n_rghost[r] += 1
            thread[r, t, 5] = first
            thread[r, t, 6] = rm
            thread[r, t, 7] = lm
    remax = max(1, max(sum(len(v) for v in e.values()) for e in rexports))
    lemax = max(1, max(sum(len(v) for v in e.values()) for e in lexports))
    rexp = np.zeros((cs, remax, 2), dtype=np.int32)
    lexp = np.zeros((cs, lemax, 2), dtype=np.int32)
    cnt = np.zeros((cs, 4), dtype=np.int32)
    for r in range(cs):
        e = 0
        for st in sorted(rexports[r]):
            lst = sorted(rexports[r][st])                # by row k: the kernels walk the list with compile-time row indices
            assert len(lst) < 256 and e < 65536
            thread[r, st, 3] = e | (len(lst) << 16)
            for (sk, dr, di) in lst:
                rexp[r, e] = (sk | (dr << 8), di)
                e += 1
        le = 0
        for st in sorted(lexports[r]):
            lst = lexports[r][st]
            assert len(lst) < 256 and le < 65536
            thread[r, st, 4] = le | (len(lst) << 16)
            for (ss, ds) in lst:
                lexp[r, le] = (ss, ds)
                le += 1
        cnt[r] = (n_rghost[r], e, le, 0)
    return StripPlan(cs, T, cpt, slots, thread, cell, rexp, lexp, cnt, segs)


def plan_for_domain(cd, cs=None, T=None, cpt=None):
    """Plan for a CompiledDomain: the requested shape, else FGB_STRIP_SHAPE, else the first of SHAPES that fits."""
    if T is not None or os.environ.get("FGB_STRIP_SHAPE"):
        return build_strip_plan(cd.sizes, cd.offsets, np.asarray(cd.nbr), cd.N, cs=cs, T=T, cpt=cpt)
    for shape in SHAPES:
        plan = build_strip_plan(cd.sizes, cd.offsets, np.asarray(cd.nbr), cd.N, cs=cs, T=shape[0], cpt=shape[1])
        if plan is not None:
            return plan
    return None


# ------------------------------------------------------------------------------------------------
# numpy emulation of the kernel's data flow (float64 or float32): used by the CPU tests to prove that the plan
# reproduces the table-driven operator and that ghost replicas stay identical to their owners.
# ------------------------------------------------------------------------------------------------
class StripEmulator:
    def __init__(self, plan: StripPlan, nbr, diag, off, dtype=np.float64):
        p = self.plan = plan
        self.dtype = dtype
        cs, T, cpt = p.cs, p.T, p.cpt
        self.real = p.cell >= 0
        self.ghost = p.cell <= -2
        kbit = (1 << np.arange(cpt))[None, None, :]
        self.rghost = (p.thread[:, :, 6:7] & kbit) != 0
        self.lghost = (p.thread[:, :, 7:8] & kbit) != 0
        assert np.array_equal(self.rghost | self.lghost, self.ghost) and not (self.rghost & self.lghost).any()
        self.gcell = np.where(self.ghost, -2 - p.cell, 0)
        self.rcell = np.where(self.real, p.cell, 0)
        self.slot = p.thread[:, :, 0:1].astype(np.int64) + np.arange(cpt)[None, None, :] * p.thread[:, :, 1:2].astype(np.int64)
        self.stride = p.thread[:, :, 1].astype(np.int64)
        # index of every remote ghost row in the receive buffer of its CTA
        self.gidx = p.thread[:, :, 5:6].astype(np.int64) + np.cumsum(self.rghost, axis=2) - self.rghost
        fm = p.thread[:, :, 2]
        self.cd = np.where(self.real, diag[self.rcell], 0).astype(dtype)
        self.co = np.zeros((4, cs, T, cpt), dtype=dtype)
        for d in range(4):
            f = (fm >> (2 * d)) & 3                                            # [cs, T]
            fb = np.broadcast_to(f[:, :, None], self.rcell.shape)
            nb = nbr[fb, self.rcell]
            self.co[d] = np.where(self.real & (nb >= 0), off[fb, self.rcell], 0)

    def scatter(self, v):
        """cell vector -> per-slot registers (real + ghost replicas), dead = 0"""
        out = np.zeros(self.plan.cell.shape, dtype=self.dtype)
        out[self.real] = v[self.rcell[self.real]]
        out[self.ghost] = v[self.gcell[self.ghost]]
        return out

    def gather(self, regs):
        out = np.zeros(int(self.rcell.max()) + 1, dtype=self.dtype)
        out[self.rcell[self.real]] = regs[self.real]
        return out

    def publish(self, regs):
        """The kernel's publish(): every thread stores its rows except its local ghost rows, then copies the rows its local
        export list names into their mirror slots (lexp); returns the per-CTA shared-memory vectors."""
        p = self.plan
        out = []
        for r in range(p.cs):
            vs = np.zeros(p.slots, dtype=self.dtype)
            keep = ~self.lghost[r]
            vs[self.slot[r][keep]] = regs[r][keep]
            n = 0
            for t in range(p.T):
                e0, c = int(p.thread[r, t, 4]) & 0xffff, int(p.thread[r, t, 4]) >> 16
                for e in range(e0, e0 + c):
                    src = int(p.lexp[r, e, 0])
                    assert src in self.slot[r, t] and not self.lghost[r, t][list(self.slot[r, t]).index(src)]   # own, real row
                    vs[int(p.lexp[r, e, 1])] = vs[src]
                    n += 1
            assert n == int(p.cnt[r, 2]) == int(self.lghost[r].sum())
            out.append(vs)
        return out

    def reload_local_ghosts(self, regs, vss):
        regs = regs.copy()
        for r in range(self.plan.cs):
            m = self.lghost[r]
            regs[r][m] = vss[r][self.slot[r][m]]
        return regs

    def spmv(self, regs, vss):
        """P v for every slot (coefficients of ghost / dead slots are zero); v of the thread's own rows from registers,
        east / west / run ends from shared memory."""
        p = self.plan
        out = np.zeros_like(regs)
        for r in range(p.cs):
            vs = vss[r]
            s = self.slot[r]
            st = self.stride[r][:, None]
            south = np.concatenate([vs[s[:, :1] - st], regs[r][:, :-1]], axis=1)
            north = np.concatenate([regs[r][:, 1:], vs[s[:, -1:] + st]], axis=1)
            out[r] = (self.cd[r] * regs[r] + self.co[0][r] * vs[s - 1] + self.co[1][r] * vs[s + 1]
                      + self.co[2][r] * south + self.co[3][r] * north)
        return out

    def push_remote(self, regs):
        """Exchange B: residuals of the mirrored rows travel along the remote export lists into the receive buffers."""
        p = self.plan
        rs = [np.zeros(max(1, int(p.cnt[r, 0])), dtype=self.dtype) for r in range(p.cs)]
        n = 0
        for r in range(p.cs):
            for t in range(p.T):
                e0, c = int(p.thread[r, t, 3]) & 0xffff, int(p.thread[r, t, 3]) >> 16
                for e in range(e0, e0 + c):
                    k, dr, di = int(p.rexp[r, e, 0]) & 0xff, int(p.rexp[r, e, 0]) >> 8, int(p.rexp[r, e, 1])
                    rs[dr][di] = regs[r, t, k]
                    n += 1
        assert n == int(self.rghost.sum()) == int(p.cnt[:, 1].sum())
        return rs

    def new_direction(self, beta, p_regs, r_regs):
        rs = self.push_remote(r_regs)
        rk = r_regs.copy()
        for r in range(self.plan.cs):
            m = self.rghost[r]
            rk[r][m] = rs[r][self.gidx[r][m]]
        p_new = beta * p_regs + rk
        vss = self.publish(p_new)
        return self.reload_local_ghosts(p_new, vss), vss

    def cg(self, f, tol, maxit, reset_steps=100):
        """The kernel's iteration (remote ghosts updated redundantly from the exchanged r, local ghosts mirrored)."""
        N = f.size
        norm = 1.0 / np.sqrt(self.dtype(N))
        fr = np.where(self.real, self.scatter(f), 0).astype(self.dtype)
        x = np.zeros(self.plan.cell.shape, dtype=self.dtype)
        r = fr.copy()
        rho = float((r * r).sum())
        p, vss = self.new_direction(0.0, np.zeros_like(x), r)
        used = -1
        for i in range(maxit):
            if reset_steps > 0 and (i + 1) % reset_steps == 0:
                r = np.where(self.real, fr - self.spmv(x, self.publish(x)), 0)
                rho = float((r * r).sum())
                p, vss = self.new_direction(0.0, p, r)
            ap = self.spmv(p, vss)
            alpha = rho / float((p * ap).sum())
            x = x + alpha * p
            r = r - alpha * ap
            rr2 = float((r * r).sum())
            used = i
            if np.sqrt(rr2) * norm < tol:
                break
            beta = rr2 / rho
            rho = rr2
            p, vss = self.new_direction(beta, p, r)
        return x, used
