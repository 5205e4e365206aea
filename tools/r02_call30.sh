# ncu launch list of the default bench command (own arm), last env.step: per-kernel shares for profiles/
set -x
O=gpurun_out/r02/final; mkdir -p $O
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $O/launches_bench_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2> $O/ncu_bench.err; tail -c 300 $O/ncu_bench.err; wc -l $O/launches_bench_final.csv
