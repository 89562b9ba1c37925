cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for C in 0 1; do
HB_RELAY_CHUNKS=$C python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$C bench.py --gpus 8 --steps 5 --warmup 3 --no-modes --no-allele --no-e2e --no-cpu > gpurun_out/bench_n8_c$C.log 2> gpurun_out/bench_n8_c$C.err
done
