cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/ngpus.txt
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 3 --no-modes > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err | cut -c1-300
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref8.log 2> gpurun_out/bench_ref8.err
tail -1 gpurun_out/bench_ref8.log | cut -c1-300
