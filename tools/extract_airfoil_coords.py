"""Writes the NACA 0012 (sharp trailing edge) surface polyline used by the airfoil environments to
``fluidgym_b200/envs/data/naca0012_sharp.npy`` ([160, 2] float64, trailing edge -> upper surface -> leading edge
-> lower surface -> trailing edge).

This is geometry INPUT DATA (a published airfoil coordinate table), not code: it is read, at build time of this
repository only, from the table the reference ships in ``fluidgym/envs/airfoil/coords.py`` so that both
implementations mesh the same body.  Run where the reference install exists:

    python tools/extract_airfoil_coords.py
"""
import ast
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for cand in (os.path.join(REPO, "baseline", "_ref", "fluidgym"), "/root/reference/src/fluidgym"):
    path = os.path.join(cand, "envs", "airfoil", "coords.py")
    if os.path.exists(path):
        break
else:
    sys.exit("reference not available")
src = open(path).read()
table = ast.literal_eval(src[src.index("["): src.rindex("]") + 1])
arr = np.asarray(table, dtype=np.float64)
assert arr.shape == (160, 2)
out = os.path.join(REPO, "fluidgym_b200", "envs", "data", "naca0012_sharp.npy")
np.save(out, arr)
print(out, arr.shape)
