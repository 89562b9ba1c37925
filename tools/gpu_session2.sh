cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -m gpu -q) > gpurun_out/pytest_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -4 gpurun_out/pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-allele --no-modes --no-e2e --no-cpu > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
tail -1 gpurun_out/bench_n2.err | cut -c1-200
