# Round-2 first GPU call (ONE GPU): everything that was committed without a GPU run, with logs kept.
set -x
mkdir -p gpurun_out/r02
for c in hooks forces cg_fused asm4; do
  timeout 600 python tests/zz_first_run_worker.py $c > gpurun_out/r02/worker_$c.log 2>&1; echo "rc=$?" >> gpurun_out/r02/worker_$c.log
done
timeout 900 python tools/extruded_check.py --json gpurun_out/r02/extruded_check.json > gpurun_out/r02/extruded_check.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -rA > gpurun_out/r02/gpu_tests.log 2>&1
timeout 600 python bench.py > gpurun_out/r02/bench.json 2> gpurun_out/r02/bench.err
FGB_ASM_ENVS=4 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02/bench_asm4.json 2> gpurun_out/r02/bench_asm4.err
FGB_ASM_ENVS=8 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02/bench_asm8.json 2> gpurun_out/r02/bench_asm8.err
timeout 900 python tools/cyl3d_bench.py --resolutions 8 24 --steps 1 --out gpurun_out/r02/cyl3d_bench.json > gpurun_out/r02/cyl3d_bench.log 2>&1
FGB_K3_CG_FUSED=1 timeout 900 python tools/cyl3d_bench.py --resolutions 8 24 --steps 1 --out gpurun_out/r02/cyl3d_bench_cg_fused.json > gpurun_out/r02/cyl3d_bench_cg_fused.log 2>&1
timeout 600 python tools/tcf_bench.py --ids TCFSmall3D-both-easy-v0 TCFLarge3D-both-easy-v0 RBC3D-easy-v0 --steps 2 --out gpurun_out/r02/tcf_bench.json > gpurun_out/r02/tcf_bench.log 2>&1
FGB_K3_CG_FUSED=1 timeout 600 python tools/tcf_bench.py --ids TCFSmall3D-both-easy-v0 TCFLarge3D-both-easy-v0 RBC3D-easy-v0 --steps 2 --out gpurun_out/r02/tcf_bench_cg_fused.json > gpurun_out/r02/tcf_bench_cg_fused.log 2>&1
tail -5 gpurun_out/r02/*.log
