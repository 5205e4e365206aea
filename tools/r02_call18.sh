set -x
O=gpurun_out/r02/exit; mkdir -p $O
timeout 200 python tools/quick_bench.py 256 11,6 > $O/quick_256.log 2>&1; grep "ms_per\|rel_l2" $O/quick_256.log
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pressure_solve or substep_matches or bit_identical" > $O/pytest.log 2>&1; tail -n 5 $O/pytest.log
