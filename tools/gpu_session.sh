cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_pooled.py -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
