set -x
bash tools/r02_tcf_sgs_grad_golden.sh
O=gpurun_out/r02/adjx3; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_extruded.py -m gpu -x -q -s -k "airfoil3d_gradients or airfoil3d_differentiable or z_invariant" > $O/pytest_a3d.log 2>&1; grep "gradients vs reference\|differentiable step\|plane \|passed\|failed\|Error" $O/pytest_a3d.log | cut -c1-600
