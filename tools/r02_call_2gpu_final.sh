# two GPUs of one box, end of round 2: slab tests (incl. the SGS viscosity halo), multi-device tests, own bench arm under torchrun
set -x
O=gpurun_out/r02/final2; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_multi_device.py -q -m gpu -rs > $O/gpu2_tests.log 2>&1
tail -8 $O/gpu2_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
tail -c 900 $O/bench_2gpu.json
tail -c 300 $O/bench_2gpu.err
