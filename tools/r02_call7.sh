set -x
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_ref_gradients.py -q -m gpu -rA > gpurun_out/r02/ref_grad_tests2.log 2>&1
tail -12 gpurun_out/r02/ref_grad_tests2.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_ref_gradients.py > gpurun_out/r02/gpu_tests_v2.log 2>&1
tail -15 gpurun_out/r02/gpu_tests_v2.log
timeout 900 python bench.py > gpurun_out/r02/bench_full.json 2> gpurun_out/r02/bench_full.err
tail -c 400 gpurun_out/r02/bench_full.err
