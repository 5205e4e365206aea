# TCF (incl. SGS) + slab tests, then the strip-kernel CTA shapes re-measured on the committed kernel (v7)
set -x
O=gpurun_out/r02/sgs; mkdir -p $O
timeout 800 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_slab.py -m gpu -q -s > $O/pytest2.log 2>&1; tail -n 25 $O/pytest2.log
for sh in 480,17 256,17 640,13 320,13 896,9; do
  FGB_STRIP_SHAPE=$sh timeout 120 python tools/quick_bench.py 256 11 > gpurun_out/r02/strip_shape_${sh/,/x}.log 2>&1; grep ms_per gpurun_out/r02/strip_shape_${sh/,/x}.log
done
FGB_GROUPS=1 timeout 120 python tools/quick_bench.py 256 11,6 > gpurun_out/r02/strip_shape_groups1.log 2>&1; grep ms_per gpurun_out/r02/strip_shape_groups1.log
