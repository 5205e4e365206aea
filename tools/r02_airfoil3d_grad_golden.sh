# gradient golden of the unmodified reference for Airfoil3D at res_z 8 (differentiable=True, one env.step = 5 solver steps)
set -x
O=gpurun_out/r02/golden3d; mkdir -p $O
timeout 2700 python oracle/ref_grad_harness.py --env Airfoil3D-easy-v0 --tag airfoil3d --res-z 8 --out $O \
   --kw '{"init_from_2d": false, "n_agents": 4}' > $O/grad_airfoil3d.log 2>&1; tail -c 300 $O/grad_airfoil3d.log
ls -la $O | tail -4
