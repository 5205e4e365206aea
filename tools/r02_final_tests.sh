# last full GPU test run of the round + smoke
set -x
O=gpurun_out/r02/final; mkdir -p $O
timeout 2400 python -m pytest tests/ -q -m gpu > $O/gpu_tests.log 2>&1; tail -n 4 $O/gpu_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/smoke.log 2>&1; tail -n 1 $O/smoke.log
