mkdir -p gpurun_out/slab
for n in 1 2 4 8; do
  for sz in "" "--large"; do
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port $((29700+n)) tools/tcf_slab_bench.py $sz --steps 20 2>/dev/null | grep '^{' | tee -a gpurun_out/slab/scale.jsonl
  done
done
