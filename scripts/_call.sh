mkdir -p gpurun_out
echo "== 7=0 pipelined kernel; 30..33 = softmax warps issue their own MMAs (EMU 3 / 2 / 4 of 8; 33 = 6-stage ring)" > gpurun_out/sweep7.log
I2V_ATTN_LIB=build/lib_exp.so timeout 300 python scripts/perf_dense_sweep.py 7=0 7=30 7=31 7=32 7=33 >> gpurun_out/sweep7.log 2>&1
SWEEP_S=9216 I2V_ATTN_LIB=build/lib_exp.so timeout 300 python scripts/perf_dense_sweep.py 7=0 7=30 >> gpurun_out/sweep7.log 2>&1
SWEEP_S=1000 I2V_ATTN_LIB=build/lib_exp.so timeout 300 python scripts/perf_dense_sweep.py 7=0 7=30 >> gpurun_out/sweep7.log 2>&1
cat gpurun_out/sweep7.log
