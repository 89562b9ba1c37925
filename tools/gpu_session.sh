cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_pooled.py tests/test_gpu_hmm.py -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
python tools/probe_pooled.py 2>&1 | tail -4 | tee gpurun_out/probe_pooled.log
