# velocity-gradient golden of the unmodified reference on the 2-D cylinder + our kernel against it
set -x
O=gpurun_out/r02/grad2d; mkdir -p $O
timeout 600 python oracle/ref_harness.py --env CylinderJet2D-easy-v0 --tag cyl24g --env-steps 1 --time-steps 0 --trace-substeps 0 --lean --gradients --out $O > $O/harness.log 2>&1; tail -n 2 $O/harness.log | cut -c1-300
python tests/golden/extract_gradient2d_fixture.py $O 2>&1 | tail -n 2
cp tests/golden/cyl24_velocity_gradients.npz $O/
rm -f $O/cyl24g_trace.npz $O/cyl24g_simstep*.npz
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "velocity_gradients" > $O/pytest.log 2>&1; grep "rel L2\|passed\|failed\|Error" $O/pytest.log | cut -c1-300
