# Airfoil3D golden of the unmodified reference at a reduced spanwise resolution (res_z 8: 46 806 x 8 = 374 448 cells)
# (gpurun brings back at most 64 MiB: the op trace and the per-sim-step states are dropped, the environment-level records kept)
set -x
O=gpurun_out/r02/airfoil3d
mkdir -p $O
timeout 2400 python oracle/ref_harness.py --env Airfoil3D-easy-v0 --tag airfoil3d --out $O --res-z 8 --env-steps 1 --time-steps 0 \
    --trace-substeps 0 --lean --kw '{"init_from_2d": false, "n_agents": 4}' > $O/airfoil3d.log 2>&1
rm -f $O/airfoil3d_trace.npz $O/airfoil3d_simstep*.npz
tail -n 3 $O/airfoil3d.log | cut -c1-400
ls -la $O
