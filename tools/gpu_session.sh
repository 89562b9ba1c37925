cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_multi.log
