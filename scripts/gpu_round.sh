#!/bin/bash
# Developer helper run under gpurun.  Parts (any subset, in order): tests bench micro sweep ncu_dense ncu_temporal
# ncu_ip ncu_ff launches sanitizer.  gpurun copies at most 64 MiB back; a full ncu capture with sources is ~20 MB.
mkdir -p gpurun_out
for part in "$@"; do
case $part in
tests)
  ( timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
  tail -5 gpurun_out/pytest_gpu.log ;;
bench)
  ( timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err )
  cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err ;;
bench_ref)
  ( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err )
  cat gpurun_out/bench_ref.log ;;
micro)
  ( ./build/pipes2; ./build/pipes; ./build/softmax_ring_step ) > gpurun_out/micro.log 2>&1; cat gpurun_out/micro.log ;;
sweep)
  timeout 600 python scripts/perf_dense_sweep.py $SWEEP_ARGS > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log ;;
ncu_dense)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_attn_pipe -s 2 -c 1 -o gpurun_out/prof_dense -f python scripts/perf_aug.py > gpurun_out/ncu_dense.log 2>&1; tail -2 gpurun_out/ncu_dense.log ;;
ncu_temporal)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:temporal_attn -s 2 -c 2 -o gpurun_out/prof_temporal -f python scripts/gpu_selftest.py --run perftemporal > gpurun_out/ncu_temporal.log 2>&1 ;;
ncu_ip)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ip_xattn -s 2 -c 1 -o gpurun_out/prof_ip -f python scripts/perf_ip_one.py > gpurun_out/ncu_ip.log 2>&1; tail -n 2 gpurun_out/ncu_ip.log ;;
ncu_ff)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ff_geglu -s 2 -c 1 -o gpurun_out/prof_ff -f python scripts/perf_ff_one.py > gpurun_out/ncu_ff.log 2>&1; tail -n 2 gpurun_out/ncu_ff.log ;;
launches)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-legs --no-e2e --eager --profiler-range > gpurun_out/bench_under_ncu.log 2>&1 ;;
sanitizer)
  ( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python scripts/sanitize_small.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log )
  tail -6 gpurun_out/sanitizer_memcheck.log
  ( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python scripts/sanitize_small.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck.log )
  tail -6 gpurun_out/sanitizer_racecheck.log
  ( timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 python scripts/sanitize_small.py > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck exit $?" >> gpurun_out/sanitizer_synccheck.log )
  tail -4 gpurun_out/sanitizer_synccheck.log ;;
*) echo "unknown part $part" ;;
esac
done
