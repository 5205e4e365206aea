# ncu evidence for the round (run under gpurun on ONE GPU): launch list of one bench step + full captures of the dominant kernels
set -x
mkdir -p gpurun_out/prof
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/prof/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cg_cluster_mb -s 8 -c 1 -f -o gpurun_out/prof/cg_push python tools/quick_bench.py 256 6 > gpurun_out/prof/ncu_cg.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3_bicgstab -s 6 -c 1 -f -o gpurun_out/prof/k3_bicg python tools/tcf_bench.py --ids TCFLarge3D-both-easy-v0 --steps 1 > gpurun_out/prof/ncu_k3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/prof/launches_tcf_large.csv python tools/tcf_bench.py --ids TCFLarge3D-both-easy-v0 --steps 1 > gpurun_out/prof/ncu_tcf.log 2>&1
ls -la gpurun_out/prof
