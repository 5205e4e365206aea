set -x
O=gpurun_out/r02/strip2; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cg_strip2 -s 6 -c 1 -f -o $O/cg_strip2_v1 python tools/quick_bench.py 74 12 > $O/ncu_v1.log 2>&1; tail -n 3 $O/ncu_v1.log
ls -la $O
