set -x
O=gpurun_out/r02/adjx3; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_extruded.py -m gpu -x -q -s -k "gradients" > $O/pytest.log 2>&1; grep -v "^$" $O/pytest.log | tail -n 30 | cut -c1-600
