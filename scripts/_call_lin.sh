mkdir -p gpurun_out
I2V_ATTN_LIB=build/lib_exp.so timeout 600 python scripts/perf_linear.py 2>&1 | tee gpurun_out/perf_linear.log
