python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tcgen05 or dilated" > gpurun_out/pytest_r1r.log 2>&1; tail -4 gpurun_out/pytest_r1r.log
for f in 2 3; do
python tools/run_attn_kernels.py 10001 6 3 $f
python tools/run_attn_kernels.py 32769 4 3 $f
done
MODALTUNE_B200_LIB=build_exp/libmt_trace.so python tools/run_attn_kernels.py 10001 1 3 2 > gpurun_out/trace_fwd2_b.log 2>&1
MODALTUNE_B200_LIB=build_exp/libmt_trace.so python tools/run_attn_kernels.py 10001 1 3 3 > gpurun_out/trace_fwd3_b.log 2>&1
