cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_expr.log 2>&1; tail -1 gpurun_out/bench_expr.log | cut -c1-1500
python bench.py --workload allele --steps 5 --warmup 3 --e2e-steps 1 > gpurun_out/bench_allele.log 2>&1; tail -1 gpurun_out/bench_allele.log | cut -c1-1500
