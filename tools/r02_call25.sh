set -x
O=gpurun_out/r02/exit; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "11" > $O/pytest_lastwarp.log 2>&1; tail -n 3 $O/pytest_lastwarp.log | cut -c1-200
timeout 200 python tools/quick_bench.py 256 11 > $O/quick_lastwarp.log 2>&1; grep "ms_per" $O/quick_lastwarp.log | cut -c1-200
timeout 200 python tools/quick_bench.py 256 11 > $O/quick_lastwarp2.log 2>&1; grep "ms_per" $O/quick_lastwarp2.log | cut -c1-200
