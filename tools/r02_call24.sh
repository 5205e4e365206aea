# two GPUs: TCFLarge slab decomposition with / without the fused BiCGStab direction update
set -x
O=gpurun_out/r02/bicgf; mkdir -p $O
for f in 0 1; do
  FGB_K3_BICG_FUSED=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port $((29720+f)) tools/tcf_slab_bench.py --large --steps 3 2>/dev/null | grep '^{' | tee $O/slab2_fused$f.json | cut -c1-400
done
FGB_K3_BICG_FUSED=1 timeout 300 python -m pytest tests/test_gpu_slab.py -q -m gpu > $O/slab2_fused_tests.log 2>&1; tail -n 2 $O/slab2_fused_tests.log
