# does the unmodified reference differentiate the z-extruded cylinder?  gradient golden at its own test resolution (res 8: 1 984 x 8 cells)
set -x
O=gpurun_out/r02/golden3d; mkdir -p $O
timeout 1700 python oracle/ref_grad_harness.py --env CylinderJet3D-easy-v0 --tag cyl3d --out $O --kw '{"resolution":8,"n_jets":8}' > $O/grad_cyl3d.log 2>&1; tail -c 400 $O/grad_cyl3d.log
ls -la $O | tail -5
