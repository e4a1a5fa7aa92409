mkdir -p gpurun_out
echo "== 7=0 default; 12 / 13 = 8- / 3-stage K/V ring; 14 / 15 = one query tile per CTA (48-key tiles), 4 / 3 CTAs per SM" > gpurun_out/sweep6.log
I2V_ATTN_LIB=build/lib_exp.so timeout 300 python scripts/perf_dense_sweep.py 7=0 7=12 7=13 7=14 7=15 >> gpurun_out/sweep6.log 2>&1
cat gpurun_out/sweep6.log
