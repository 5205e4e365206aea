"""CylinderJet3D / extruded D = 3 launch path on the GPU (tools/extruded_check.py: substep, reset and env.step against the
unmodified reference's goldens).  When this file was committed the round's GPU budget was spent and the launch path had NEVER run
on a GPU -- everything around it is verified on the CPU (tests/test_cylinder3d_cpu.py, test_extruded_host.py).  The check therefore
runs LAST (file name), in its OWN PROCESS with a time limit, so that a fault in the new path cannot disturb the verified suites,
and it is a non-strict xfail: XPASS = the path is verified on this box, XFAIL = see the JSON lines it prints."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.xfail(strict=False, reason="first GPU run of the extruded launch path (never executed on a GPU when committed)")
def test_extruded_path_first_gpu_run(tmp_path):
    out = tmp_path / "extruded_check.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "extruded_check.py"), "--json", str(out)], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:], file=sys.stderr)
    try:                                                       # keep the evidence where gpurun brings it back
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "extruded_check.log"), "w") as f:
            f.write(r.stdout + "\n--- stderr ---\n" + r.stderr)
    except OSError:
        pass
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    verdict = json.load(open(out))
    assert verdict["ok"]
