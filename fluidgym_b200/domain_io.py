"""Reader / writer of the reference's on-disk domain format (``simulation/pict/util/domain_io.py:64-327``): ``<path>.json``
describes the block / boundary structure and refers to flat tensors stored under integer keys in ``<path>.npz``.  This is
the format of the published initial-domain splits (``initial_domains/<env id>/<idx>/<mode>.{json,npz}``,
``envs/fluid_env.py:1044-1112``), so states written by the reference load into the batched environments and vice versa.
"""
from __future__ import annotations

import json

import numpy as np

from .domain import CONNECTED, FIXED, PERIODIC, Boundary, DomainSpec


def _scalar_type(bd) -> str:
    """``passiveScalarType`` of a boundary entry: the reference writes a list (one entry per channel, domain_io.py:145-148) and
    reads a list or a plain string (:291-302)."""
    t = bd.get("passiveScalarType", "DIRICHLET")
    t = t[0] if isinstance(t, (list, tuple)) else t
    return str(t)


def load_domain(path: str):
    """-> (DomainSpec, state) with state = dict(u [2,N], p [N], bvel [2,NB], T [N] | None, sbval [NB] | None) in the flat
    layout of the solver (cells of all blocks concatenated, x fastest; prescribed faces in (block, face) order)."""
    with open(path + ".json") as fh:
        d = json.load(fh)
    with np.load(path + ".npz") as z:
        data = {k: z[k] for k in z.files}

    def get(dct, name):
        return data[dct[name]] if name in dct else None
    if d["spatialDims"] != 2:
        raise NotImplementedError("domain files of 3-D domains: use fluidgym_b200.box3d (only 2-D multi-block domains are read here)")
    n_scalar = d.get("passiveScalarChannels", 1)           # the reference's loader defaults to one channel (domain_io.py:207-209)
    svisc = get(d, "passiveScalarViscosity")
    spec = DomainSpec(float(np.asarray(get(d, "viscosity")).ravel()[0]), d.get("name", "domain"),
                      scalar_viscosity=None if (n_scalar == 0 or svisc is None) else float(np.asarray(svisc).ravel()[0]))
    for blk in d["blocks"]:
        v = get(blk, "vertexCoordinates")
        if v is None:
            raise NotImplementedError("blocks stored by transform only (no vertex coordinates)")
        spec.create_block(v[0], blk.get("name", ""))
    us, ps, Ts = [], [], []
    for bi, blk in enumerate(d["blocks"]):
        b = spec.blocks[bi]
        us.append(get(blk, "velocity")[0].reshape(2, -1))
        ps.append(get(blk, "pressure")[0].reshape(-1))
        if "scalar" in blk and n_scalar:
            Ts.append(get(blk, "scalar")[0].reshape(-1))
        for f, bd in enumerate(blk["boundaries"]):
            kind = bd["type"]
            if kind == "FIXED":
                if bd.get("velocityType", "DIRICHLET") != "DIRICHLET":
                    raise NotImplementedError(f"boundary velocity type {bd.get('velocityType')}")
                vel = np.asarray(get(bd, "velocity"), dtype=np.float32)
                vel = vel.reshape(2, -1) if vel.ndim > 2 else vel.reshape(2, 1)
                sc, neumann = None, False
                if "scalar" in bd and n_scalar:
                    sc = np.asarray(get(bd, "scalar"), dtype=np.float32).reshape(-1)
                    neumann = _scalar_type(bd) == "NEUMANN"
                # the file lists every face explicitly: set it directly (close_boundary would also re-close the partner face)
                n = b.size(1 - (f >> 1))
                vfull = np.zeros((2, n), dtype=np.float32)
                vfull[:] = vel
                sfull = np.zeros(n, dtype=np.float32)
                if sc is not None:
                    sfull[:] = sc
                b.bounds[f] = Boundary(FIXED, velocity=vfull, scalar=sfull, scalar_neumann=neumann)
            elif kind == "CONNECTED":
                other, (of, axis) = int(bd["connectedBlock"]), bd["axes"]
                b.bounds[f] = Boundary(CONNECTED, other, (int(of), int(axis)))
            elif kind == "PERIODIC":
                b.bounds[f] = Boundary(PERIODIC)
            else:
                raise NotImplementedError(f"boundary type {kind}")
    bvel, sbval = [], []
    for bi, b in enumerate(spec.blocks):
        for f in range(4):
            if b.bounds[f].type == FIXED:
                bvel.append(np.asarray(b.bounds[f].velocity, dtype=np.float32).reshape(2, -1))
                sbval.append(np.asarray(b.bounds[f].scalar, dtype=np.float32).reshape(-1))
    state = dict(u=np.concatenate(us, axis=1).astype(np.float32), p=np.concatenate(ps).astype(np.float32),
                 bvel=np.concatenate(bvel, axis=1) if bvel else np.zeros((2, 0), np.float32),
                 T=np.concatenate(Ts).astype(np.float32) if Ts else None,
                 sbval=np.concatenate(sbval) if (sbval and n_scalar) else None)
    return spec, state


def save_domain(spec: DomainSpec, state: dict, path: str):
    """Write ``spec`` with the given flat state in the reference's format (load_domain of either implementation reads it)."""
    data = []

    def add(arr, dct, name):
        dct[name] = str(len(data))
        data.append(np.ascontiguousarray(arr, dtype=np.float32))
    has_T = state.get("T") is not None and spec.scalar_viscosity is not None
    d = {"name": spec.name, "spatialDims": 2}
    add(np.array([spec.viscosity]), d, "viscosity")
    d["passiveScalarChannels"] = 1 if has_T else 0
    if has_T:
        add(np.array([spec.scalar_viscosity]), d, "passiveScalarViscosity")
    d["blocks"] = []
    o = ob = 0
    for bi, b in enumerate(spec.blocks):
        n = b.nx * b.ny
        bd = {"name": b.name}
        add(state["u"][:, o:o + n].reshape(1, 2, b.ny, b.nx), bd, "velocity")
        add(state["p"][o:o + n].reshape(1, 1, b.ny, b.nx), bd, "pressure")
        if has_T:
            add(state["T"][o:o + n].reshape(1, 1, b.ny, b.nx), bd, "scalar")
        add(b.vertex[None], bd, "vertexCoordinates")
        o += n
        bd["boundaries"] = []
        for f in range(4):
            bn = b.bounds[f]
            if bn.type == FIXED:
                m = b.size(1 - (f >> 1))
                e = {"type": "FIXED", "velocityType": "DIRICHLET"}
                shape = (1, 2, m, 1) if (f >> 1) == 0 else (1, 2, 1, m)
                add(state["bvel"][:, ob:ob + m].reshape(shape), e, "velocity")
                if has_T:
                    e["passiveScalarType"] = ["NEUMANN" if bn.scalar_neumann else "DIRICHLET"]
                    add(state["sbval"][ob:ob + m].reshape((1, 1) + shape[2:]), e, "scalar")
                ob += m
            elif bn.type == CONNECTED:
                e = {"type": "CONNECTED", "connectedBlock": int(bn.other), "axes": [int(bn.axes[0]), int(bn.axes[1])]}
            else:
                e = {"type": "PERIODIC"}
            bd["boundaries"].append(e)
        d["blocks"].append(bd)
    d["data_info"] = {str(i): {"shape": list(a.shape), "dtype": "float32", "device": "cpu"} for i, a in enumerate(data)}
    np.savez_compressed(path + ".npz", **{str(i): a for i, a in enumerate(data)})
    with open(path + ".json", "w") as fh:
        json.dump(d, fh)


# ---------------------------------------------------------------------------------------------------------------------
# D = 3: single-block rectilinear boxes (TCF, RBC3D).  Same file format; tensors carry one more spatial axis:
# velocity [1,3,z,y,x], pressure / scalar [1,1,z,y,x], vertexCoordinates [1,3,z+1,y+1,x+1]; a FIXED boundary stores its
# velocity as [1,3] (static) or with a singleton on the face-normal axis, e.g. [1,3,z,1,x] on a y face.
# ---------------------------------------------------------------------------------------------------------------------
def load_box_domain(path: str):
    """-> dict(vertex [3,nz+1,ny+1,nx+1], closed (x,y,z), viscosity, scalar_viscosity | None, name,
    state = dict(u [3,N], p [N], bvel [3,NB], T [N] | None, sbval [NB] | None)) in the layout of fluidgym_b200.box3d
    (cell g = x + nx (y + ny z); prescribed faces face-major (-x, +x, -y, ...), tangential cells in (z, y, x) order)."""
    with open(path + ".json") as fh:
        d = json.load(fh)
    with np.load(path + ".npz") as z:
        data = {k: z[k] for k in z.files}

    def get(dct, name):
        return data[dct[name]] if name in dct else None
    if d["spatialDims"] != 3 or len(d["blocks"]) != 1:
        raise NotImplementedError("load_box_domain reads 3-D single-block domains (2-D multi-block domains: load_domain)")
    blk = d["blocks"][0]
    v = get(blk, "vertexCoordinates")
    if v is None:
        raise NotImplementedError("blocks stored by transform only (no vertex coordinates)")
    vertex = np.ascontiguousarray(v[0], dtype=np.float32)
    nz, ny, nx = (s - 1 for s in vertex.shape[1:])
    N = nx * ny * nz
    n_scalar = d.get("passiveScalarChannels", 1)           # the reference's loader defaults to one channel (domain_io.py:207-209)
    svisc = get(d, "passiveScalarViscosity")
    u = np.asarray(get(blk, "velocity"), dtype=np.float32)[0].reshape(3, N)
    p = np.asarray(get(blk, "pressure"), dtype=np.float32)[0].reshape(N)
    T = np.asarray(get(blk, "scalar"), dtype=np.float32)[0].reshape(N) if ("scalar" in blk and n_scalar) else None
    closed = [False, False, False]
    bvel, sbval = [], []
    for f, bd in enumerate(blk["boundaries"]):
        kind, ax = bd["type"], f >> 1
        if kind == "PERIODIC":
            continue
        if kind != "FIXED":
            raise NotImplementedError(f"boundary type {kind} on a single-block box")
        if bd.get("velocityType", "DIRICHLET") != "DIRICHLET":
            raise NotImplementedError(f"boundary velocity type {bd.get('velocityType')}")
        closed[ax] = True
        n_face = N // (nx, ny, nz)[ax]
        vel = np.asarray(get(bd, "velocity"), dtype=np.float32)
        vfull = np.zeros((3, n_face), dtype=np.float32)
        vfull[:] = vel.reshape(3, -1) if vel.ndim > 2 else vel.reshape(3, 1)
        bvel.append(vfull)
        sfull = np.zeros(n_face, dtype=np.float32)
        if "scalar" in bd and n_scalar:
            if _scalar_type(bd) != "DIRICHLET":
                raise NotImplementedError("Neumann scalar boundaries on the 3-D box")
            sfull[:] = np.asarray(get(bd, "scalar"), dtype=np.float32).reshape(-1)
        sbval.append(sfull)
    for ax in range(3):
        kinds = {blk["boundaries"][2 * ax]["type"], blk["boundaries"][2 * ax + 1]["type"]}
        if len(kinds) != 1:
            raise NotImplementedError("an axis of the box must be periodic or closed on both sides")
    state = dict(u=u, p=p, bvel=np.concatenate(bvel, axis=1) if bvel else np.zeros((3, 0), np.float32), T=T,
                 sbval=np.concatenate(sbval) if (sbval and n_scalar) else None)
    return dict(vertex=vertex, closed=tuple(closed), viscosity=float(np.asarray(get(d, "viscosity")).ravel()[0]),
                scalar_viscosity=None if (n_scalar == 0 or svisc is None) else float(np.asarray(svisc).ravel()[0]),
                name=d.get("name", "domain"), state=state)


def save_box_domain(vertex, closed, viscosity: float, state: dict, path: str, scalar_viscosity=None, name="domain"):
    """Write a 3-D single-block box with the given flat state in the reference's format."""
    vertex = np.ascontiguousarray(vertex, dtype=np.float32)
    nz, ny, nx = (s - 1 for s in vertex.shape[1:])
    data = []

    def add(arr, dct, key):
        dct[key] = str(len(data))
        data.append(np.ascontiguousarray(arr, dtype=np.float32))
    has_T = state.get("T") is not None and scalar_viscosity is not None
    d = {"name": name, "spatialDims": 3}
    add(np.array([viscosity]), d, "viscosity")
    d["passiveScalarChannels"] = 1 if has_T else 0
    if has_T:
        add(np.array([scalar_viscosity]), d, "passiveScalarViscosity")
    bd = {"name": "block"}
    add(np.asarray(state["u"]).reshape(1, 3, nz, ny, nx), bd, "velocity")
    add(np.asarray(state["p"]).reshape(1, 1, nz, ny, nx), bd, "pressure")
    if has_T:
        add(np.asarray(state["T"]).reshape(1, 1, nz, ny, nx), bd, "scalar")
    add(vertex[None], bd, "vertexCoordinates")
    bd["boundaries"] = []
    ob = 0
    for f in range(6):
        ax = f >> 1
        if not closed[ax]:
            bd["boundaries"].append({"type": "PERIODIC"})
            continue
        shape = [nz, ny, nx]
        shape[2 - ax] = 1
        m = shape[0] * shape[1] * shape[2]
        e = {"type": "FIXED", "velocityType": "DIRICHLET"}
        add(np.asarray(state["bvel"])[:, ob:ob + m].reshape([1, 3] + shape), e, "velocity")
        if has_T:
            e["passiveScalarType"] = ["DIRICHLET"]
            add(np.asarray(state["sbval"])[ob:ob + m].reshape([1, 1] + shape), e, "scalar")
        ob += m
        bd["boundaries"].append(e)
    d["blocks"] = [bd]
    d["data_info"] = {str(i): {"shape": list(a.shape), "dtype": "float32", "device": "cpu"} for i, a in enumerate(data)}
    np.savez_compressed(path + ".npz", **{str(i): a for i, a in enumerate(data)})
    with open(path + ".json", "w") as fh:
        json.dump(d, fh)


# ---------------------------------------------------------------------------------------------------------------------
# D = 3: z-extruded multi-block domains (CylinderJet3D, Airfoil3D).  The reference's writer is generic over the dimension
# (util/domain_io.py:64-122): tensors carry one more spatial axis as for the boxes above, every block lists six boundaries (the
# z pair PERIODIC), and a CONNECTED boundary stores ``axes = [otherFace, axis1, axis2]`` where axis1 / axis2 are the faces of the
# other block that the block's axes (face/2 + 1) % 3 and (face/2 + 2) % 3 are aligned with (envs/cylinder/grid.py:376-415): for an
# x face the in-plane tangential axis y comes first and z second, for a y face z comes first and x second.  Read into the layout
# of fluidgym_b200/extruded3d.py: the 2-D DomainSpec of the plane, u [3, nz, N2], p [nz, N2], bvel [3, nz, NB2].
# PARITY NOTE: the 2-D multi-block and 3-D box formats are pinned to files written by the reference; no reference-written file of a
# 3-D multi-block domain exists in this repository yet, so the axes triple above follows the reference's source, not a fixture.
# ---------------------------------------------------------------------------------------------------------------------
def load_extruded_domain(path: str):
    """-> (DomainSpec of the plane, z_vertices [nz+1], state = dict(u [3,nz,N2], p [nz,N2], bvel [3,nz,NB2]))"""
    with open(path + ".json") as fh:
        d = json.load(fh)
    with np.load(path + ".npz") as z:
        data = {k: z[k] for k in z.files}

    def get(dct, name):
        return data[dct[name]] if name in dct else None
    if d["spatialDims"] != 3:
        raise NotImplementedError("load_extruded_domain reads 3-D multi-block domains (2-D: load_domain)")
    if d.get("passiveScalarChannels", 0):
        raise NotImplementedError("passive scalars on extruded multi-block domains")
    spec = DomainSpec(float(np.asarray(get(d, "viscosity")).ravel()[0]), d.get("name", "domain"))
    zv = None
    for blk in d["blocks"]:
        v = get(blk, "vertexCoordinates")
        if v is None:
            raise NotImplementedError("blocks stored by transform only (no vertex coordinates)")
        v = np.asarray(v[0], dtype=np.float32)                                     # [3, nz+1, ny+1, nx+1]
        if not (np.array_equal(v[:2], np.broadcast_to(v[:2, :1], v[:2].shape)) and np.array_equal(v[2], np.broadcast_to(v[2, :, :1, :1], v[2].shape))):
            raise NotImplementedError("the block is not a z-extrusion of a 2-D grid")
        bz = v[2, :, 0, 0]
        if zv is None:
            zv = bz
        elif not np.array_equal(zv, bz):
            raise NotImplementedError("blocks with different z planes")
        spec.create_block(np.ascontiguousarray(v[:2, 0]), blk.get("name", ""))
    nz = zv.size - 1
    us, ps = [], []
    for bi, blk in enumerate(d["blocks"]):
        b = spec.blocks[bi]
        us.append(np.asarray(get(blk, "velocity"), dtype=np.float32)[0].reshape(3, nz, -1))
        ps.append(np.asarray(get(blk, "pressure"), dtype=np.float32)[0].reshape(nz, -1))
        if len(blk["boundaries"]) != 6 or any(blk["boundaries"][f]["type"] != "PERIODIC" for f in (4, 5)):
            raise NotImplementedError("extruded blocks must be periodic in z")
        for f in range(4):
            bd = blk["boundaries"][f]
            kind = bd["type"]
            if kind == "FIXED":
                if bd.get("velocityType", "DIRICHLET") != "DIRICHLET":
                    raise NotImplementedError(f"boundary velocity type {bd.get('velocityType')}")
                n = b.size(1 - (f >> 1))
                vel = np.asarray(get(bd, "velocity"), dtype=np.float32)
                vfull = np.zeros((3, nz, n), dtype=np.float32)
                vfull[:] = vel.reshape(3, nz, n) if vel.ndim > 2 else vel.reshape(3, 1, 1)
                bnd = Boundary(FIXED, velocity=np.ascontiguousarray(vfull[:2, 0]), scalar=np.zeros(n, dtype=np.float32))
                bnd.velocity3 = vfull
                b.bounds[f] = bnd
            elif kind == "CONNECTED":
                of, a1, a2 = (int(x) for x in bd["axes"])
                tang, zal = (a1, a2) if (f >> 1) == 0 else (a2, a1)
                if zal != 4:
                    raise NotImplementedError("block connections that flip or permute the z axis")
                b.bounds[f] = Boundary(CONNECTED, int(bd["connectedBlock"]), (of, tang))
            elif kind == "PERIODIC":
                b.bounds[f] = Boundary(PERIODIC)
            else:
                raise NotImplementedError(f"boundary type {kind}")
    bvel = [b.bounds[f].velocity3 for b in spec.blocks for f in range(4) if b.bounds[f].type == FIXED]
    state = dict(u=np.concatenate(us, axis=2), p=np.concatenate(ps, axis=1),
                 bvel=np.concatenate(bvel, axis=2) if bvel else np.zeros((3, nz, 0), np.float32))
    return spec, np.asarray(zv, dtype=np.float32), state


def save_extruded_domain(spec: DomainSpec, z_vertices, state: dict, path: str):
    """Write the z-extrusion of ``spec`` over ``z_vertices`` with state u [3,nz,N2], p [nz,N2], bvel [3,nz,NB2] in the reference's
    format (see the parity note above)."""
    zv = np.asarray(z_vertices, dtype=np.float32)
    nz = zv.size - 1
    data = []

    def add(arr, dct, name):
        dct[name] = str(len(data))
        data.append(np.ascontiguousarray(arr, dtype=np.float32))
    d = {"name": spec.name, "spatialDims": 3}
    add(np.array([spec.viscosity]), d, "viscosity")
    d["passiveScalarChannels"] = 0
    d["blocks"] = []
    u, p, bv = (np.asarray(state[k], dtype=np.float32) for k in ("u", "p", "bvel"))
    o = ob = 0
    for b in spec.blocks:
        n = b.nx * b.ny
        bd = {"name": b.name}
        add(u[:, :, o:o + n].reshape(1, 3, nz, b.ny, b.nx), bd, "velocity")
        add(p[:, o:o + n].reshape(1, 1, nz, b.ny, b.nx), bd, "pressure")
        v3 = np.zeros((1, 3, nz + 1, b.ny + 1, b.nx + 1), dtype=np.float32)
        v3[0, :2] = b.vertex[:, None]
        v3[0, 2] = zv[:, None, None]
        add(v3, bd, "vertexCoordinates")
        o += n
        bd["boundaries"] = []
        for f in range(4):
            bn = b.bounds[f]
            if bn.type == FIXED:
                m = b.size(1 - (f >> 1))
                e = {"type": "FIXED", "velocityType": "DIRICHLET"}
                shape = (1, 3, nz, m, 1) if (f >> 1) == 0 else (1, 3, nz, 1, m)
                add(bv[:, :, ob:ob + m].reshape(shape), e, "velocity")
                ob += m
            elif bn.type == CONNECTED:
                of, tang = int(bn.axes[0]), int(bn.axes[1])
                e = {"type": "CONNECTED", "connectedBlock": int(bn.other), "axes": [of, tang, 4] if (f >> 1) == 0 else [of, 4, tang]}
            else:
                e = {"type": "PERIODIC"}
            bd["boundaries"].append(e)
        bd["boundaries"] += [{"type": "PERIODIC"}, {"type": "PERIODIC"}]
        d["blocks"].append(bd)
    d["data_info"] = {str(i): {"shape": list(a.shape), "dtype": "float32", "device": "cpu"} for i, a in enumerate(data)}
    np.savez_compressed(path + ".npz", **{str(i): a for i, a in enumerate(data)})
    with open(path + ".json", "w") as fh:
        json.dump(d, fh)
