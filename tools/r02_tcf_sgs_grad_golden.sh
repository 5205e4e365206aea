# gradient golden of the unmodified reference for the channel WITH the Smagorinsky model + our test against it
set -x
O=gpurun_out/r02/golden3d; mkdir -p $O
timeout 1200 python oracle/ref_grad_harness.py --env TCFSmall3D-both-easy-v0 --tag tcf32_sgs --perturb 0.05 --out $O \
   --kw '{"resolution_x_z":32,"resolution_y":33,"init_with_noise":false,"C_smag":0.1,"use_van_driest":true}' > $O/grad_tcf32_sgs.log 2>&1; tail -c 200 $O/grad_tcf32_sgs.log
python tests/golden/extract_grad_fixtures.py $O tcf_sgs 2>&1 | tail -c 600
cp tests/golden/tcf32_sgs_grad.npz $O/tcf32_sgs_grad_fixture.npz
timeout 600 python -m pytest tests/test_gpu_tcf.py -m gpu -x -q -s -k "gradients" > $O/pytest_tcf_sgs.log 2>&1; grep "gradients vs reference\|passed\|failed" $O/pytest_tcf_sgs.log | cut -c1-500
