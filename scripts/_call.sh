mkdir -p gpurun_out
echo "== experiments lib (no-exp pipeline floors: 7=5 one thread per row, 7=6 two)" > gpurun_out/sweep3.log
I2V_ATTN_LIB=build/lib_exp.so timeout 300 python scripts/perf_dense_sweep.py 7=0 7=5 7=6 >> gpurun_out/sweep3.log 2>&1
cat gpurun_out/sweep3.log
I2V_ATTN_LIB=build/lib_exp.so timeout 600 python -m pytest tests/test_gpu_linear.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/linear_test.log
I2V_ATTN_LIB=build/lib_exp.so timeout 600 python scripts/perf_linear.py 2>&1 | tee gpurun_out/perf_linear.log
bash scripts/gpu_round.sh tests
