# D = 3 Krylov kernels: merged rho reduction (5 grid-wide exchanges per BiCGStab iteration) + structured neighbour arithmetic
set -x
O=gpurun_out/r02/k3; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_slab.py tests/test_gpu_rbc3d.py -m gpu -x -q > $O/pytest.log 2>&1; tail -n 8 $O/pytest.log
timeout 600 python tools/tcf_bench.py --ids TCFSmall3D-both-easy-v0 TCFLarge3D-both-easy-v0 RBC3D-easy-v0 --steps 2 --out $O/tcf_bench_box.json > $O/tcf_bench_box.log 2>&1; tail -n 4 $O/tcf_bench_box.log
FGB_O3_BOX=0 timeout 600 python tools/tcf_bench.py --ids TCFSmall3D-both-easy-v0 TCFLarge3D-both-easy-v0 RBC3D-easy-v0 --steps 2 --out $O/tcf_bench_table.json > $O/tcf_bench_table.log 2>&1; tail -n 4 $O/tcf_bench_table.log
