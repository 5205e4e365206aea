set -x
mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "11" > gpurun_out/r02/strip_tests.log 2>&1
tail -5 gpurun_out/r02/strip_tests.log
for shape in 480,17 256,17; do
FGB_STRIP_SHAPE=$shape timeout 120 python tools/quick_bench.py 256 11 > gpurun_out/r02/strip_quick_$shape.log 2>&1
cat gpurun_out/r02/strip_quick_$shape.log
done
FGB_STRIP_SHAPE=480,17 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_cg_strip -s 8 -c 1 -f -o gpurun_out/r02/cg_strip_v7_480 python tools/quick_bench.py 256 11 > gpurun_out/r02/ncu_strip_v7.log 2>&1
