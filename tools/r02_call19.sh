set -x
O=gpurun_out/r02/adj3; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tcf.py -m gpu -x -q -s -k "gradients_match or substep_matches or structured" > $O/pytest.log 2>&1; grep -v "^$" $O/pytest.log | tail -n 25 | cut -c1-400
