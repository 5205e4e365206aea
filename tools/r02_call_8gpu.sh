# eight GPUs of one box: the own bench arm under torchrun with the extras (RBC weak scaling, TCFLarge slab strong-scaling point)
set -x
O=gpurun_out/r02/final8; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv | head -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 2 --warmup 3 --extra-airfoil-steps 1 > $O/bench_8gpu.json 2> $O/bench_8gpu.err
python -c "
import json;d=json.loads([l for l in open("gpurun_out/r02/final8/bench_8gpu.json") if l.startswith("{")][-1]);print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, {k:(v.get('value'),v.get('ms_per_substep'),v.get('n_gpus'),v.get('mode')) for k,v in d['extra'].items()})"
tail -c 300 $O/bench_8gpu.err
