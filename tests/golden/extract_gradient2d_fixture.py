#!/usr/bin/env python
"""Reduce ``oracle/ref_harness.py --env CylinderJet2D-easy-v0 --tag cyl24g --env-steps 1 --time-steps 0 --trace-substeps 0 --lean
--gradients`` (the UNMODIFIED reference; tools/r02_grad2d_golden.sh) to ``tests/golden/cyl24_velocity_gradients.npz``: the state after
the env.step (velocity, boundary velocities) and PISOtorch.ComputeSpatialVelocityGradients of it, in the flat layout of the product
(cells of all blocks concatenated); grad [component c][direction d][N] -- the reference's list index is the component, the tensor
channel the direction.  Test infrastructure only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
from extract_fixtures import _multiblock_helpers  # noqa: E402


def main(src, tag="cyl24g"):
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    nb, faces, cells, bfaces = _multiblock_helpers(make_cylinder_domain(24))
    st, gr = np.load(os.path.join(src, f"{tag}_state_step0.npz")), np.load(os.path.join(src, f"{tag}_gradients.npz"))
    grad = np.stack([np.concatenate([gr[f"b{bi}_d{c}"].reshape(2, -1) for bi in range(nb)], axis=1) for c in range(2)]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "cyl24_velocity_gradients.npz"), u=cells(st, "b{}_u", 2), bvel=bfaces(st, "b{}_f{}_velocity"), grad=grad)
    print("cyl24_velocity_gradients.npz", grad.shape, float(np.abs(grad).max()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "..", "gpurun_out", "r02", "grad2d"))
