MODALTUNE_B200_LIB=build_exp/libmt_trace.so python tools/run_attn_kernels.py 10001 1 3 2 > gpurun_out/trace_fwd2_a.log 2>&1
MODALTUNE_B200_LIB=build_exp/libmt_trace.so python tools/run_attn_kernels.py 10001 1 3 3 > gpurun_out/trace_fwd3_a.log 2>&1
grep -c "smx\|mma " gpurun_out/trace_fwd2_a.log gpurun_out/trace_fwd3_a.log
