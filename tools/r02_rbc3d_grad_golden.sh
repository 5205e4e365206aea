# gradient golden of the unmodified reference for RBC3D (16 x 10 x 16 cells, 2 x 2 heaters of 8 cells; with 4-cell heaters the
# reference's own action gradient is NaN) + our reverse-mode test against it
set -x
O=gpurun_out/r02/golden3d; mkdir -p $O
timeout 1200 python oracle/ref_grad_harness.py --env RBC3D-easy-v0 --tag rbc3d --seed 1 --out $O \
   --kw '{"n_heaters":2,"resolution":8,"step_length":0.25,"use_marl":false}' > $O/grad_rbc3d.log 2>&1; tail -c 300 $O/grad_rbc3d.log
python tests/golden/extract_grad_fixtures.py $O rbc3d 2>&1 | tail -c 1500
cp tests/golden/rbc3d_grad.npz $O/rbc3d_grad_fixture.npz
timeout 600 python -m pytest tests/test_gpu_rbc3d.py -m gpu -x -q -s -k "gradients" > $O/pytest_rbc3d.log 2>&1; grep -v "^$" $O/pytest_rbc3d.log | tail -n 25 | cut -c1-600
