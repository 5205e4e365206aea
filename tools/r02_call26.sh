set -x
O=gpurun_out/r02/bicgf; mkdir -p $O
timeout 600 python tools/tcf_bench.py --ids TCFSmall3D-both-easy-v0 TCFLarge3D-both-easy-v0 RBC3D-easy-v0 --steps 2 --out $O/tcf_bench_batch4.json > $O/tcf_bench_batch4.log 2>&1; grep -o '"env": "[^"]*"\|"ms_per_substep": [0-9.]*' $O/tcf_bench_batch4.log | paste - -
timeout 300 python -m pytest tests/test_gpu_tcf.py -m gpu -x -q -k "substep_matches or structured or fused_bicgstab" > $O/pytest3.log 2>&1; tail -n 2 $O/pytest3.log | cut -c1-200
