set -x
O=gpurun_out/r02/fuse; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_slab.py tests/test_gpu_rbc3d.py -m gpu -x -q > $O/pytest.log 2>&1; tail -n 3 $O/pytest.log | cut -c1-300
timeout 600 python tools/tcf_bench.py --ids TCFSmall3D-both-easy-v0 TCFLarge3D-both-easy-v0 RBC3D-easy-v0 --steps 2 --out $O/tcf_bench.json > $O/tcf_bench.log 2>&1; grep -o '"env": "[^"]*"\|"ms_per_substep": [0-9.]*\|"launches_per_substep": [0-9.]*' $O/tcf_bench.log | paste - - -
