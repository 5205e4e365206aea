set -x
O=gpurun_out/r02/final; mkdir -p $O
timeout 600 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/one_line_a.json 2> $O/one_line_a.err; wc -l $O/one_line_a.json; python -c "import json;d=json.loads(open('$O/one_line_a.json').read());print(d['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 1 --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/one_line_b.json 2> $O/one_line_b.err; wc -l $O/one_line_b.json; python -c "import json;d=json.loads(open('$O/one_line_b.json').read());print(d['value'])"; grep -c "NCCL version" $O/one_line_b.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 --ref-envs 1 > $O/one_line_c.json 2> $O/one_line_c.err; wc -l $O/one_line_c.json; python -c "import json;d=json.loads(open('$O/one_line_c.json').read());print(d['value'], d['impl'])"
