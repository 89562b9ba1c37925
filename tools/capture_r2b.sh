# ncu --set full captures of the kernels the second half of round 2 rebuilt; summaries only come back
# (the .ncu-rep files together exceed what gpurun copies home).  Run under gpurun on one GPU.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out /tmp/rep
N="ncu --set full --clock-control none --import-source on"
$N -k regex:pool_accum_cm -c 1 -o /tmp/rep/pool0 python tools/probe_pool_mean.py > /dev/null 2>&1
$N -k regex:pool_accum_cm -s 1 -c 1 -o /tmp/rep/pool1 python tools/probe_pool_mean.py > /dev/null 2>&1
$N -k regex:dist_sq_pipe -c 1 -o /tmp/rep/dist_clean python tools/probe_dist.py > /dev/null 2>&1
$N -k regex:dist_sq_pipe -s 6 -c 1 -o /tmp/rep/dist_masked python tools/probe_dist.py > /dev/null 2>&1
python tools/summarize_ncu.py /tmp/rep/pool0.ncu-rep gpurun_out/r2b_pool_accum_pass0_full.txt 1000000000 > /dev/null
python tools/summarize_ncu.py /tmp/rep/pool1.ncu-rep gpurun_out/r2b_pool_accum_pass1_full.txt 1000000000 > /dev/null
python tools/summarize_ncu.py /tmp/rep/dist_clean.ncu-rep gpurun_out/r2b_dist_pipe_clean_full.txt > /dev/null
python tools/summarize_ncu.py /tmp/rep/dist_masked.ncu-rep gpurun_out/r2b_dist_pipe_masked_full.txt > /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_poolmean.csv python tools/probe_pool_mean.py > /dev/null 2>&1
ls -la gpurun_out | tail -8
