mkdir -p gpurun_out/slab
for n in 1 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port $((29800+n)) tools/tcf_slab_bench.py --large --nz-mult 8 --steps 10 --warmup 3 2>/dev/null | grep '^{' | tee -a gpurun_out/slab/scale_big.jsonl
done
