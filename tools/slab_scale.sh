# strong scaling of one channel domain over the GPUs of a node (run under: gpurun --gpus 8 -- bash tools/slab_scale.sh)
mkdir -p gpurun_out/slab
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port $((29700+$1)) tools/tcf_slab_bench.py "${@:2}" 2>/dev/null | grep '^{' | tee -a gpurun_out/slab/scale_r2.jsonl; }
run 1 --large --plain --steps 20
for n in 1 2 4 8; do run $n --large --steps 20; done
run 1 --large --nz-mult 8 --plain --steps 10 --warmup 3
run 8 --large --nz-mult 8 --steps 10 --warmup 3
