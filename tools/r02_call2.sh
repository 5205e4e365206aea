set -x
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "11 or 6-" > gpurun_out/r02/strip_tests.log 2>&1
tail -15 gpurun_out/r02/strip_tests.log
timeout 600 python tools/quick_bench.py 256 6,11 > gpurun_out/r02/strip_quick.log 2>&1
cat gpurun_out/r02/strip_quick.log
