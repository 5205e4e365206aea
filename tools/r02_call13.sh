# cg_impl 12 (two environments per cluster, interleaved): parity tests first (short timeouts: a protocol bug would hang), then timing
set -x
O=gpurun_out/r02/strip2; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "12 or sim_step_and_outflow" > $O/pytest.log 2>&1; tail -n 15 $O/pytest.log
timeout 150 python tools/quick_bench.py 256 12,11 > $O/quick_256.log 2>&1; grep "ms_per\|rel_l2" $O/quick_256.log
FGB_GROUPS=1 timeout 150 python tools/quick_bench.py 256 12,11 > $O/quick_256_g1.log 2>&1; grep "ms_per\|rel_l2" $O/quick_256_g1.log
timeout 150 python tools/quick_bench.py 37 12,11 > $O/quick_37.log 2>&1; grep "ms_per\|rel_l2" $O/quick_37.log
