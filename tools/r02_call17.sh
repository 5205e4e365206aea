set -x
O=gpurun_out/r02/k3; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_slab.py tests/test_gpu_rbc3d.py -m gpu -x -q > $O/pytest2.log 2>&1; tail -n 8 $O/pytest2.log
timeout 600 python tools/tcf_bench.py --ids TCFSmall3D-both-easy-v0 TCFLarge3D-both-easy-v0 RBC3D-easy-v0 --steps 2 --out $O/tcf_bench_box2.json > $O/tcf_bench_box2.log 2>&1; tail -n 4 $O/tcf_bench_box2.log
