# First gpurun call of the next round (ONE GPU): the checks that could not run when round 1's GPU budget was spent.
#   /usr/local/graft/bin/gpurun --timeout 3600 -- 'bash tools/r02_first_call.sh'
# 1. first GPU run of the extruded D = 3 launch path + CylinderJet3D environment against the reference's goldens
# 2. the whole GPU suite
# 3. the default bench line
# 4. reference goldens for Airfoil3D at a small resolution (op trace + env.step), for the next extruded environment
set -x
mkdir -p gpurun_out/r02
timeout 900 python tools/extruded_check.py --json gpurun_out/r02/extruded_check.json > gpurun_out/r02/extruded_check.log 2>&1
FGB_X3_HOOKS=cuda timeout 900 python tools/extruded_check.py --json gpurun_out/r02/extruded_check_cuda_hooks.json > gpurun_out/r02/extruded_check_cuda_hooks.log 2>&1
timeout 1200 python tools/cyl3d_bench.py --resolutions 8 24 --steps 1 --airfoil3d --out gpurun_out/r02/cyl3d_bench.json > gpurun_out/r02/cyl3d_bench.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02/gpu_tests.log 2>&1
timeout 600 python bench.py > gpurun_out/r02/bench.json 2> gpurun_out/r02/bench.err
FGB_ASM_ENVS=4 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02/bench_asm4.json 2> gpurun_out/r02/bench_asm4.err
FGB_ASM_ENVS=8 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02/bench_asm8.json 2> gpurun_out/r02/bench_asm8.err
# A/B of the fused-update cooperative CG (bit-identical, one grid.sync less per iteration)
FGB_K3_CG_FUSED=1 timeout 900 python tools/cyl3d_bench.py --resolutions 8 24 --steps 1 --out gpurun_out/r02/cyl3d_bench_cg_fused.json > gpurun_out/r02/cyl3d_bench_cg_fused.log 2>&1
timeout 600 python tools/tcf_bench.py --ids TCFSmall3D-both-easy-v0 RBC3D-easy-v0 --steps 2 --out gpurun_out/r02/tcf_bench.json > gpurun_out/r02/tcf_bench.log 2>&1
FGB_K3_CG_FUSED=1 timeout 600 python tools/tcf_bench.py --ids TCFSmall3D-both-easy-v0 RBC3D-easy-v0 --steps 2 --out gpurun_out/r02/tcf_bench_cg_fused.json > gpurun_out/r02/tcf_bench_cg_fused.log 2>&1
# launch list + one full capture of the pressure CG of the extruded path (CylinderJet3D res 24), for profiles/
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02/launches_cyl3d_res24.csv \
    python tools/cyl3d_bench.py --resolutions 24 --steps 1 --settle 0 > gpurun_out/r02/ncu_cyl3d_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_cg -s 6 -c 1 -f -o gpurun_out/r02/k3_cg_cyl3d \
    python tools/cyl3d_bench.py --resolutions 24 --steps 1 --settle 0 > gpurun_out/r02/ncu_cyl3d_cg.log 2>&1
# (res_z = 96: 46 806 x 96 = 4.5 M cells, the reference needs minutes per env.step -> lean trace, one env.step, generous limit)
timeout 1700 python oracle/ref_harness.py --env Airfoil3D-easy-v0 --tag airfoil3d --out gpurun_out/r02/airfoil3d --env-steps 1 --time-steps 0 \
    --trace-substeps 1 --lean --kw '{"init_from_2d": false}' > gpurun_out/r02/airfoil3d.log 2>&1
# a reference-written 3-D multi-block domain file, to pin fluidgym_b200/domain_io.py::load_extruded_domain
timeout 600 python oracle/ref_harness.py --env CylinderJet3D-easy-v0 --tag cyl3d --out gpurun_out/r02/dom_cyl3d --save-domain-only --env-steps 0 \
    --kw '{"resolution": 8, "n_jets": 8}' > gpurun_out/r02/dom_cyl3d.log 2>&1
tail -3 gpurun_out/r02/*.log
