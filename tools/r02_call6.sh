set -x
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_ref_gradients.py -q -m gpu -rA > gpurun_out/r02/ref_grad_tests.log 2>&1
tail -30 gpurun_out/r02/ref_grad_tests.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "groups" > gpurun_out/r02/groups_tests.log 2>&1
tail -5 gpurun_out/r02/groups_tests.log
timeout 300 python tests/zz_first_run_worker.py hooks > gpurun_out/r02/worker_hooks2.log 2>&1; echo "rc=$?" >> gpurun_out/r02/worker_hooks2.log
tail -5 gpurun_out/r02/worker_hooks2.log
for g in 1 2 4; do
FGB_GROUPS=$g timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r02/bench_groups$g.json 2> gpurun_out/r02/bench_groups$g.err
done
FGB_GROUPS=2 FGB_CG_IMPL=11 timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r02/bench_groups2_impl11.json 2> gpurun_out/r02/bench_groups2_impl11.err
FGB_GROUPS=4 FGB_CG_IMPL=11 timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r02/bench_groups4_impl11.json 2> gpurun_out/r02/bench_groups4_impl11.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 2 > gpurun_out/r02/bench_reference.json 2> gpurun_out/r02/bench_reference.err
tail -c 600 gpurun_out/r02/bench_reference.err
