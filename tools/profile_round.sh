set -x
mkdir -p gpurun_out/prof
python bench.py --steps 8 --warmup 3 > gpurun_out/prof/bench_cylinder.json 2> gpurun_out/prof/bench_cylinder.err
python bench.py --workload rbc --envs 1024 --steps 4 --warmup 3 > gpurun_out/prof/bench_rbc.json 2> gpurun_out/prof/bench_rbc.err
python tools/quick_bench.py 256 1,2,3 > gpurun_out/prof/quick_impls.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/prof/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cg_cluster_mb -s 8 -c 1 -f -o gpurun_out/prof/cg_mb python tools/quick_bench.py 256 3 > gpurun_out/prof/ncu_cg.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bicgstab_cluster -s 4 -c 1 -f -o gpurun_out/prof/bicg python tools/quick_bench.py 256 3 > gpurun_out/prof/ncu_bicg.log 2>&1
ls -la gpurun_out/prof
cat gpurun_out/prof/bench_cylinder.json gpurun_out/prof/bench_rbc.json gpurun_out/prof/quick_impls.txt
