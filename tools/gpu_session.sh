cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
ncu --set full --clock-control none --import-source on -k regex:binom2_fast -s 3 -c 1 -o gpurun_out/prof_binom2_fast_v1 -f python tools/probe_binom.py honeybadger_b200/libhb_b200.so > gpurun_out/ncu_b2f.log 2>&1
tail -2 gpurun_out/ncu_b2f.log
