cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_r1b.log 2>&1; tail -1 gpurun_out/bench_r1b.log
