set -x
O=gpurun_out/r02/k3; mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_tcf_large.csv python tools/tcf_bench.py --ids TCFLarge3D-both-easy-v0 --steps 1 > $O/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k3_bicgstab -s 6 -c 1 -f -o $O/k3_bicgstab_r02 python tools/tcf_bench.py --ids TCFLarge3D-both-easy-v0 --steps 1 > $O/ncu_k3.log 2>&1; tail -n 2 $O/ncu_k3.log
ls -la $O
