# full GPU suite + smoke + default bench (own arm), logs under gpurun_out/r02/final
set -x
O=gpurun_out/r02/final; mkdir -p $O
timeout 2400 python -m pytest tests/ -q -m gpu > $O/gpu_tests.log 2>&1; tail -n 15 $O/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 1500 $O/bench.json
