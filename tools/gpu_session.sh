cd $GRAFT_REPO_ROOT
HB_MODE=0 python tools/probe_binom.py variants/lib_p2.so variants/lib_p4.so variants/lib_p2b5.so variants/lib_p2b3.so 2>&1 | grep "^lib" | tee gpurun_out/probe_binom3.log
