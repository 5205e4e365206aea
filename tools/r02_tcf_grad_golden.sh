# does the unmodified reference differentiate the 3-D channel?  gradient golden on the 32 x 33 x 32 grid (one env.step)
set -x
O=gpurun_out/r02/golden3d; mkdir -p $O
timeout 1200 python oracle/ref_grad_harness.py --env TCFSmall3D-both-easy-v0 --tag tcf32 --perturb 0.05 --out $O \
   --kw '{"resolution_x_z":32,"resolution_y":33,"init_with_noise":false}' > $O/grad_tcf32.log 2>&1; tail -n 12 $O/grad_tcf32.log
ls -la $O
