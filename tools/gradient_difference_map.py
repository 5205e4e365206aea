"""Where does our d reward / d u0 of the 2-D cylinder differ from the reference's?  CPU only.

Inputs: the reference's gradient goldens (tests/golden/cyl24[_tight]_grad.npz, written by the unmodified reference with
differentiable=True) and ours from the same state (written by tests/test_gpu_ref_gradients.py on a B200, kept as
profiles/r02_ours_cyl24[_tight]_grads.npz).  Output: share of the squared difference per 3x3 neighbourhood of the eight points
where a block connection ends on a prescribed boundary (markdown, profiles/r02_cyl24_gradient_difference_map.md)."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain  # noqa: E402

R = 3
cd = make_cylinder_domain(24).prepare()
offs, sizes = np.array(cd.offsets), cd.sizes


def cells(bi, xs, ys):
    nx, ny = sizes[bi]
    return [int(offs[bi] + (y % ny) * nx + (x % nx)) for y in ys for x in xs]


lo, hi = range(0, R), range(-R, 0)
# blocks: 0 left (inflow -x, cylinder +x), 1 top (cylinder -y, wall +y), 2 right (cylinder -x), 3 bottom (wall -y, cylinder +y), 4 wake
POINTS = {
    "outer, top-left: left / top blocks, inflow + wall": cells(0, lo, hi) + cells(1, lo, hi),
    "outer, bottom-left: left / bottom blocks, inflow + wall": cells(0, lo, lo) + cells(3, lo, lo),
    "outer, top-right: right / top / wake blocks, wall": cells(2, hi, hi) + cells(1, hi, hi) + cells(4, lo, hi),
    "outer, bottom-right: right / bottom / wake blocks, wall": cells(2, hi, lo) + cells(3, hi, lo) + cells(4, lo, lo),
    "cylinder, left / top": cells(0, hi, hi) + cells(1, lo, lo),
    "cylinder, left / bottom": cells(0, hi, lo) + cells(3, lo, hi),
    "cylinder, right / top": cells(2, lo, hi) + cells(1, hi, lo),
    "cylinder, right / bottom": cells(2, lo, lo) + cells(3, hi, hi),
}


def main():
    out = []
    for tag, label in (("cyl24", "reference at its default tolerances"), ("cyl24_tight", "both codes at 1e-7 tolerances")):
        ours = np.load(os.path.join(ROOT, "profiles", f"r02_ours_{tag}_grads.npz"))
        ref = np.load(os.path.join(ROOT, "tests", "golden", f"{tag}_grad.npz"))
        for key, what in (("dreward_du", "d reward / d u0"), ("vjp_du", "state vjp (reference's cotangent on the outgoing velocity)")):
            d, g = ours[key] - ref[key], ref[key]
            tot = (d ** 2).sum()
            out.append(f"\n**{what}, {label}**: relative L2 difference {np.sqrt(tot) / np.linalg.norm(g):.2e}\n")
            out.append("| 3x3 cells per block around | share of the squared difference | local relative difference | share of the gradient's norm |")
            out.append("|---|---|---|---|")
            seen = set()
            for name, c in POINTS.items():
                c = np.array(c)
                seen |= set(c.tolist())
                out.append("| %s | %.3f | %.1e | %.3f |" % (name, (d[:, c] ** 2).sum() / tot, np.linalg.norm(d[:, c]) / np.linalg.norm(g[:, c]),
                                                       np.linalg.norm(g[:, c]) / np.linalg.norm(g)))
            rest = np.array(sorted(set(range(cd.N)) - seen))
            out.append("| all other %d cells | %.3f | %.1e | %.3f |" % (len(rest), (d[:, rest] ** 2).sum() / tot,
                                                                     np.linalg.norm(d[:, rest]) / np.linalg.norm(g[:, rest]),
                                                                     np.linalg.norm(g[:, rest]) / np.linalg.norm(g)))
    print("\n".join(out))


if __name__ == "__main__":
    main()
