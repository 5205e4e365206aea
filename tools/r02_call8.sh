set -x
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_ref_gradients.py -q -m gpu > gpurun_out/r02/ref_grad_tests5.log 2>&1
tail -8 gpurun_out/r02/ref_grad_tests5.log
timeout 900 python -m pytest tests/test_gpu_autograd.py tests/test_gpu_rbc_autograd.py tests/test_gpu_airfoil.py -q -m gpu > gpurun_out/r02/autograd_tests.log 2>&1
tail -6 gpurun_out/r02/autograd_tests.log
