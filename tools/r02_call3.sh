set -x
mkdir -p gpurun_out/r02
FGB_STRIP_SHAPE=640,13 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cg_strip -s 8 -c 1 -f -o gpurun_out/r02/cg_strip_v3_640 python tools/quick_bench.py 256 11 > gpurun_out/r02/ncu_strip_v3.log 2>&1
FGB_STRIP_SHAPE=160,13 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cg_strip -s 8 -c 1 -f -o gpurun_out/r02/cg_strip_v3_160 python tools/quick_bench.py 256 11 > gpurun_out/r02/ncu_strip_v3b.log 2>&1
