set -x
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests/test_gpu_extruded.py -q -m gpu -k airfoil3d -s > gpurun_out/r02/airfoil3d_test.log 2>&1
tail -12 gpurun_out/r02/airfoil3d_test.log | cut -c1-900
