#!/bin/bash
# Developer helper run under gpurun: GPU tests, bench, ncu launch list and full captures of the hot kernels.
# gpurun copies at most 64 MiB back, and a full capture with sources is ~19 MB: part 1 = tests, bench, launch list, dense +
# temporal captures; part 2 = IP-Adapter + feed-forward GEMM captures.
mkdir -p gpurun_out
part=${1:-1}
if [ "$part" = "1" ]; then
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
( timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err )
( timeout 300 python bench.py --steps 5 --warmup 3 --eager --no-cpu-baseline > gpurun_out/bench_eager.log 2>> gpurun_out/bench.err )
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --eager --profiler-range > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_attn_pipe -s 2 -c 1 -o gpurun_out/prof_dense -f python scripts/perf_aug.py > gpurun_out/ncu_dense.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:temporal_attn -s 2 -c 2 -o gpurun_out/prof_temporal -f python scripts/gpu_selftest.py --run perftemporal > gpurun_out/ncu_temporal.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
else
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ip_xattn -s 2 -c 1 -o gpurun_out/prof_ip -f python scripts/perf_ip_one.py > gpurun_out/ncu_ip.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ff_geglu -s 2 -c 1 -o gpurun_out/prof_ff -f python scripts/perf_ff_one.py > gpurun_out/ncu_ff.log 2>&1
tail -n 2 gpurun_out/ncu_ip.log; tail -n 2 gpurun_out/ncu_ff.log
fi
