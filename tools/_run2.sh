python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/pytest_r1q.log 2>&1; tail -4 gpurun_out/pytest_r1q.log
for f in 2 3; do
python tools/run_attn_kernels.py 10001 6 3 $f
python tools/run_attn_kernels.py 32769 4 3 $f
done
python tools/profile_step.py > gpurun_out/profile_step_r1q.log 2>&1; grep -E "wall|total|gated|cross_" gpurun_out/profile_step_r1q.log | cut -c1-150
