set -x
O=gpurun_out/r02/adj3; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tcf.py -m gpu -x -q -s -k "tcf_gradients" > $O/pytest_b2.log 2>&1; grep "gradients vs reference\|passed\|failed\|Error\|assert" $O/pytest_b2.log | cut -c1-400
