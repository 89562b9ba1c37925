cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for F in 0 1; do
HB_E2E_FULL_PLANES=$F python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2955$F bench.py --gpus 8 --steps 5 --warmup 3 --no-modes --no-allele --no-sharded --no-cpu --e2e-steps 3 > gpurun_out/bench_n8_f$F.log 2> gpurun_out/bench_n8_f$F.err
done
nproc > gpurun_out/nproc8.txt; free -g | head -2 >> gpurun_out/nproc8.txt
