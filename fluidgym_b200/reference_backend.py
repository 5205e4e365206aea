"""The drop-in seam, executed: ``B200Simulation`` stands in for ``fluidgym.simulation.Simulation`` at ``single_step()``
inside an UNMODIFIED reference environment (last row).

The reference's environments own a mutable C++ ``PISOtorch.Domain`` (blocks with velocity / pressure tensors, ``FixedBoundary``
objects the actuators write with ``setVelocity``) and call ``self._sim.single_step()`` once per solver step
(simulation/simulation.py:210-280; envs/cylinder/cylinder_env_base.py:741-776).  Everything around that call -- actuation, force
integration (envs/util/forces.py), sensor rendering -- reads and writes the Domain through the pybind11 accessors of
simulation/extensions/PISOtorch.cpp:99-504.  This module keeps all of that untouched:

    pull   Domain -> flat SoA state        block.velocity / domain.pressureResult / FixedBoundary.velocity      (device to device)
    step   fgb_sim_step                    adaptive CFL substeps, outflow relaxation + flux balance, PISO substeps (sm_100a kernels)
    push   flat SoA state -> Domain        block.setVelocity / setPressure, domain.setVelocityResult / setPressureResult,
                                           FixedBoundary.setVelocity on the advective-outflow faces, domain.UpdateDomainData()

so the reference's own ``env.step`` computes its reward and observations from fields advanced by this library.  A maintainer's patch
is two lines in ``CylinderEnvBase._get_simulation`` (cylinder_env_base.py:303-332), see ``patch_reference_env`` below and
INTEGRATION.md section 1; tests/test_gpu_reference_backend.py runs it against the reference's stock run.

Nothing here imports the reference: it only talks to the objects it is handed (duck typing on the pybind11 API).
"""
from __future__ import annotations

import numpy as np
import torch

from .domain import CONNECTED, FIXED, PERIODIC, Boundary, CompiledDomain, DomainSpec
from .solver import BatchedPISO


def domain_to_spec(domain):
    """``PISOtorch.Domain`` (2-D multi-block) -> (DomainSpec, cell transforms, boundary transforms), using only public accessors:
    ``getBlocks``, ``Block.vertexCoordinates / transform / getBoundary``, ``FixedBoundary.velocity / transform``,
    ``ConnectedBoundary.getConnectedBlock / axes`` (PISOtorch.cpp:99-330).  The reference's own metric tensors are passed on, so
    the compiled tables use bit-identical geometry."""
    if domain.getSpatialDims() != 2:
        raise NotImplementedError("reference_backend: 2-D multi-block domains (3-D boxes: fluidgym_b200.box3d)")
    blocks = list(domain.getBlocks())
    spec = DomainSpec(float(domain.viscosity.detach().cpu().reshape(-1)[0]), "reference-domain")
    transforms, btransforms = [], {}
    for blk in blocks:
        spec.create_block(blk.vertexCoordinates[0].detach().cpu().numpy())
        transforms.append(blk.transform.detach().cpu().numpy().reshape(-1, 9))
    for bi, blk in enumerate(blocks):
        b = spec.blocks[bi]
        for f in range(4):
            bnd = blk.getBoundary(f)
            kind = type(bnd).__name__
            if kind == "FixedBoundary":
                n = b.size(1 - (f >> 1))
                vel = np.zeros((2, n), dtype=np.float32)
                vel[:] = bnd.velocity.detach().cpu().numpy().reshape(2, -1)
                b.bounds[f] = Boundary(FIXED, velocity=vel, scalar=np.zeros(n, dtype=np.float32), scalar_neumann=False)
                if bnd.hasTransform():
                    btransforms[(bi, f)] = bnd.transform.detach().cpu().numpy().reshape(-1, 9)
            elif kind == "ConnectedBoundary":
                other = next(i for i, o in enumerate(blocks) if o is bnd.getConnectedBlock() or o == bnd.getConnectedBlock())
                axes = list(bnd.axes)
                b.bounds[f] = Boundary(CONNECTED, other, (int(axes[0]), int(axes[1])))
            elif kind == "PeriodicBoundary":
                b.bounds[f] = Boundary(PERIODIC)
            else:
                raise NotImplementedError(f"reference_backend: boundary type {kind}")
    return spec, transforms, btransforms


class B200Simulation:
    """``Simulation.single_step()`` of the reference on the sm_100a solver, state living in the reference's ``Domain``.

    ``stock``: the reference's own Simulation object; every attribute this class does not define (``output_resampling_*``,
    ``time_step`` ...) resolves there, so the environment's rendering / observation code keeps working.
    ``outflow``: [(block index, face 0..3)] of the advective-outflow boundaries the stock "PRE" prep function updates
    (cylinder_env_base.py:277-300), ``char_vel`` its advection velocity, ``bc_tol`` its flux-balance tolerance."""

    def __init__(self, stock, domain, dt, adaptive_cfl=0.8, outflow=(), char_vel=(1.0, 0.0), bc_tol=5e-6, **solver_kw):
        self._stock, self._domain = stock, domain
        self._dt, self._cfl, self._char_vel, self._bc_tol = float(dt), float(adaptive_cfl), tuple(char_vel), float(bc_tol)
        spec, transforms, btransforms = domain_to_spec(domain)
        self.spec = spec
        self.cd = cd = CompiledDomain(spec, transforms=transforms, btransforms=btransforms or None)
        self._blocks = list(domain.getBlocks())
        self._faces = [(bi, f) for bi, b in enumerate(spec.blocks) for f in range(4) if b.bounds[f].type == FIXED]
        out_mask = np.zeros(cd.NB, dtype=bool)
        self._out_faces = []
        for bi, f in outflow:
            o, n = int(cd.boff[bi, f]), spec.blocks[bi].size(1 - (f >> 1))
            out_mask[o:o + n] = True
            self._out_faces.append((bi, f, o, n))
        dev = self._blocks[0].velocity.device
        self.solver = BatchedPISO(cd, 1, device=str(dev), out_mask=out_mask if outflow else None, **solver_kw)
        self.substeps = 0

    def __getattr__(self, name):              # only called for attributes not defined here
        return getattr(self._stock, name)

    # ---- Domain <-> flat state ------------------------------------------------------------------------------------------
    def pull(self):
        s, dom = self.solver, self._domain
        s.u[0].copy_(torch.cat([b.velocity.detach().reshape(2, -1) for b in self._blocks], dim=1))
        s.p[0].copy_(dom.pressureResult.detach().reshape(-1))
        s.buffer("ures")[0].copy_(dom.velocityResult.detach().reshape(2, -1))
        parts = []
        for bi, f in self._faces:
            n = self.spec.blocks[bi].size(1 - (f >> 1))
            parts.append(self._blocks[bi].getBoundary(f).velocity.detach().reshape(2, -1).expand(2, n))
        if parts:
            s.bvel[0].copy_(torch.cat(parts, dim=1))

    def push(self):
        s, dom = self.solver, self._domain
        off = 0
        for blk in self._blocks:
            shp = blk.velocity.shape
            n = shp[-1] * shp[-2]
            blk.setVelocity(s.u[0, :, off:off + n].reshape(shp).contiguous().clone())
            blk.setPressure(s.p[0, off:off + n].reshape(blk.pressure.shape).contiguous().clone())
            off += n
        dom.setVelocityResult(s.u[0].reshape(dom.velocityResult.shape).contiguous().clone())
        dom.setPressureResult(s.p[0].reshape(dom.pressureResult.shape).contiguous().clone())
        for bi, f, o, n in self._out_faces:
            bnd = self._blocks[bi].getBoundary(f)
            v = s.bvel[0, :, o:o + n]
            shape = list(bnd.velocity.shape)
            if bnd.velocity.numel() != 2 * n:                     # static boundary value: make room for one value per face
                shape = [1, 2, n, 1] if (f >> 1) == 0 else [1, 2, 1, n]
            bnd.setVelocity(v.reshape(shape).contiguous().clone())
        dom.UpdateDomainData()

    # ---- the seam ---------------------------------------------------------------------------------------------------------
    def single_step(self, static: bool = False) -> bool:
        """simulation/simulation.py:210-280 (adaptive branch): one solver step of length dt."""
        if static:
            return self._stock.single_step(static=True)
        self.pull()
        self.substeps += self.solver.single_step(self._dt, self._cfl, char_vel=self._char_vel if self._out_faces else None,
                                                 bc_tol=self._bc_tol)
        self.push()
        return True


def patch_reference_env(env_base_cls, outflow, char_vel=(1.0, 0.0), bc_tol=5e-6, **solver_kw):
    """Monkey-patch ``env_base_cls._get_simulation`` (e.g. fluidgym.envs.cylinder.cylinder_env_base.CylinderEnvBase) so that every
    environment built afterwards steps on this library while everything else stays the reference's code.  Returns the
    original method (to undo the patch)."""
    orig = env_base_cls._get_simulation

    def _get_simulation(self, domain, prep_fn):
        stock = orig(self, domain, prep_fn)                                   # incl. its make_divergence_free
        return B200Simulation(stock, domain, dt=self._dt, adaptive_cfl=self._adaptive_cfl, outflow=outflow, char_vel=char_vel,
                              bc_tol=bc_tol, **solver_kw)

    env_base_cls._get_simulation = _get_simulation
    return orig
