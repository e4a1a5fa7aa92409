mkdir -p gpurun_out
N=${1:-2}
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "exit $?" >> gpurun_out/bench_n$N.err )
tail -c 600 gpurun_out/bench_n$N.log; tail -4 gpurun_out/bench_n$N.err
