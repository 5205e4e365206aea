set -x
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_reference_backend.py tests/test_gpu_airfoil.py tests/test_gpu_parity.py -q -m gpu -k "backend or airfoil_differentiable or env_step_matches" -s > gpurun_out/r02/backend_test.log 2>&1
tail -25 gpurun_out/r02/backend_test.log
