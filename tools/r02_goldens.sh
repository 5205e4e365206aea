# Reference-generated goldens of round 2 (ONE GPU, the unmodified reference in baseline/_ref):
#  (a) gradients through env.step with differentiable=True (cylinder, RBC, airfoil)
#  (b) tight-tolerance / fp64 traces of one substep (cylinder, RBC): measures the "CG tolerance ball" of DESIGN section 5
set -x
O=gpurun_out/r02/golden
mkdir -p $O
timeout 900 python oracle/ref_grad_harness.py --env CylinderJet2D-easy-v0 --tag cyl24 --out $O > $O/grad_cyl24.log 2>&1
timeout 900 python oracle/ref_grad_harness.py --env CylinderJet2D-easy-v0 --tag cyl24_tight --pressure-tol 1e-7 --advection-tol 1e-7 --out $O > $O/grad_cyl24_tight.log 2>&1
timeout 900 python oracle/ref_grad_harness.py --env RBC2D-easy-v0 --tag rbc --out $O > $O/grad_rbc.log 2>&1
timeout 1500 python oracle/ref_grad_harness.py --env Airfoil2D-medium-v0 --tag airfoil --out $O > $O/grad_airfoil.log 2>&1
timeout 600 python oracle/ref_harness.py --env CylinderJet2D-easy-v0 --tag cyl24_f64 --dtype float64 --pressure-tol 1e-11 --advection-tol 1e-11 --env-steps 1 --time-steps 0 --trace-substeps 2 --lean --out $O > $O/trace_cyl24_f64.log 2>&1
timeout 600 python oracle/ref_harness.py --env CylinderJet2D-easy-v0 --tag cyl24_t32 --pressure-tol 1e-7 --advection-tol 1e-7 --env-steps 1 --time-steps 0 --trace-substeps 2 --lean --out $O > $O/trace_cyl24_t32.log 2>&1
timeout 600 python oracle/ref_harness.py --env RBC2D-easy-v0 --tag rbc_f64 --dtype float64 --pressure-tol 1e-11 --advection-tol 1e-11 --env-steps 1 --time-steps 0 --trace-substeps 2 --lean --out $O > $O/trace_rbc_f64.log 2>&1
for f in $O/*.log; do echo "== $f"; tail -n 4 $f; done
ls -la $O
