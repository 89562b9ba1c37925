cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2>&1; tail -1 gpurun_out/bench_n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['cpu_baseline'])"
