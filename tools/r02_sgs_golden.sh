# Smagorinsky golden of the unmodified reference (baseline/_ref) on a 32 x 33 x 32 channel + our parity tests against it, one GPU
set -x
O=gpurun_out/r02/sgs
mkdir -p $O/fixtures
timeout 900 python oracle/ref_harness.py --env TCFSmall3D-both-easy-v0 --tag tcf32_sgs --perturb 0.05 --env-steps 1 --time-steps 0 \
    --trace-substeps 1 --lean --gradients --out $O \
    --kw '{"resolution_x_z":32,"resolution_y":33,"init_with_noise":false,"C_smag":0.1,"use_van_driest":true}' > $O/harness.log 2>&1
tail -n 5 $O/harness.log
python tests/golden/extract_tcf_fixtures.py $O > $O/extract.log 2>&1; tail -n 6 $O/extract.log
cp tests/golden/tcf32_sgs_* $O/fixtures/
rm -f $O/tcf32_sgs_trace.npz $O/tcf32_sgs_state_*.npz $O/tcf32_sgs_simstep*.npz $O/tcf32_sgs_geometry.npz
timeout 900 python -m pytest tests/test_gpu_tcf.py -m gpu -x -q -s > $O/pytest.log 2>&1; tail -n 30 $O/pytest.log
ls -la $O $O/fixtures
