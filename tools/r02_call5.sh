set -x
mkdir -p gpurun_out/r02
FGB_CG_IMPL=11 timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_zz_gpu_extruded_first_run.py > gpurun_out/r02/gpu_tests_impl11.log 2>&1
tail -15 gpurun_out/r02/gpu_tests_impl11.log
timeout 300 python bench.py --workload rbc --envs 1024 --no-cpu-baseline > gpurun_out/r02/bench_rbc_6.json 2> gpurun_out/r02/bench_rbc_6.err
FGB_CG_IMPL=11 timeout 300 python bench.py --workload rbc --envs 1024 --no-cpu-baseline > gpurun_out/r02/bench_rbc_11.json 2> gpurun_out/r02/bench_rbc_11.err
timeout 300 python tools/airfoil_bench.py --envs 1 8 32 --diff-steps 1 > gpurun_out/r02/airfoil_6.log 2>&1
FGB_CG_IMPL=11 timeout 300 python tools/airfoil_bench.py --envs 1 8 32 --diff-steps 1 > gpurun_out/r02/airfoil_11.log 2>&1
tail -3 gpurun_out/r02/airfoil_6.log gpurun_out/r02/airfoil_11.log
