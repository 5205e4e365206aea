#!/bin/bash
# end-of-round check after the LinsolveError watch: full GPU suite, smoke, a short bench line (airfoil extra cut to one env.step)
set -x
O=gpurun_out/r02/last
mkdir -p $O
timeout 420 python -m pytest tests/ -q -m gpu > $O/gpu_tests.log 2>&1
tail -n 6 $O/gpu_tests.log
timeout 120 python -c 'import __graft_entry__ as g; g.smoke(); print("SMOKE_OK")' > $O/smoke.log 2>&1
tail -n 1 $O/smoke.log
timeout 300 python bench.py --no-cpu-baseline --extra-airfoil-steps 1 > $O/bench.json 2> $O/bench.err
wc -l $O/bench.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02/last/bench.json").read())
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], list(d["extra"]))
PY
