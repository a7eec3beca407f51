python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1s.log 2>&1; tail -2 gpurun_out/pytest_r1s.log
python tools/profile_step.py > gpurun_out/profile_step_r1s.log 2>&1; grep -E "wall|total|gated|cross_|dilated" gpurun_out/profile_step_r1s.log | cut -c1-150
python bench.py > gpurun_out/bench_r1s.log 2>&1; tail -1 gpurun_out/bench_r1s.log | cut -c1-200
