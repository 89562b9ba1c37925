cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_expr_r1.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_allele_r1.csv python bench.py --workload allele --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:norm3_fast -s 3 -c 1 -o gpurun_out/prof_norm3_fast_r1 -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:binom2_fast -s 3 -c 1 -o gpurun_out/prof_binom2_fast_r1 -f python bench.py --workload allele --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_f2.log 2>&1
tail -2 gpurun_out/ncu_f2.log
