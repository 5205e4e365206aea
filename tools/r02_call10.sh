set -x
mkdir -p gpurun_out/r02
FGB_STRIP_SHAPE=320,13 timeout 120 python tools/quick_bench.py 256 11 > gpurun_out/r02/strip_quick_320b.log 2>&1
cat gpurun_out/r02/strip_quick_320b.log
FGB_STRIP_SHAPE=320,13 timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r02/bench_shape320.json 2> gpurun_out/r02/bench_shape320.err
FGB_GROUPS=8 timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r02/bench_groups8.json 2> gpurun_out/r02/bench_groups8.err
for g in 1 2 4; do
FGB_GROUPS=$g timeout 300 python bench.py --workload rbc --envs 1024 --no-cpu-baseline --no-extras > gpurun_out/r02/bench_rbc_groups$g.json 2> gpurun_out/r02/bench_rbc_groups$g.err
done
python - <<'PY'
import json
for f in ["bench_shape320","bench_groups8","bench_rbc_groups1","bench_rbc_groups2","bench_rbc_groups4"]:
    try:
        d=json.loads(open(f"gpurun_out/r02/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],1), round(d["e2e"]["value"]))
    except Exception as e: print(f, "ERR", e)
PY
