cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_hmm.py -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
HB_MODE=0 python tools/probe_binom.py variants/lib_r4.so variants/lib_r5.so variants/lib_r6.so 2>&1 | grep "^lib" | tee gpurun_out/probe_binom5.log
