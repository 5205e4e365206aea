# full GPU suite + smoke + default bench (own arm) + a short reference-arm line, logs under gpurun_out/r02/final
set -x
O=gpurun_out/r02/final; mkdir -p $O
timeout 2400 python -m pytest tests/ -q -m gpu > $O/gpu_tests.log 2>&1; tail -n 6 $O/gpu_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/smoke.log 2>&1; tail -n 1 $O/smoke.log
timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; python -c "
import json;d=json.load(open('$O/bench.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['extra']['tcf_large']['ms_per_substep'], d['clocks'])"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 600 $O/bench_reference.json
