# two GPUs of one box: slab-decomposition parity test, both bench arms under torchrun
set -x
mkdir -p gpurun_out/r02
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_gpu_slab.py -q -m gpu -rs > gpurun_out/r02/slab_2gpu_tests.log 2>&1
tail -8 gpurun_out/r02/slab_2gpu_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02/bench_2gpu.json 2> gpurun_out/r02/bench_2gpu.err
tail -c 1500 gpurun_out/r02/bench_2gpu.json
tail -c 600 gpurun_out/r02/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 2 --ref-envs 1 > gpurun_out/r02/bench_ref_2gpu.json 2> gpurun_out/r02/bench_ref_2gpu.err
tail -c 600 gpurun_out/r02/bench_ref_2gpu.json
