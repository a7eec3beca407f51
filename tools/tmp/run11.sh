python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1u.log 2>&1; tail -3 gpurun_out/pytest_r1u.log
python bench.py > gpurun_out/bench_r1u.log 2>&1; tail -1 gpurun_out/bench_r1u.log | cut -c1-220
python tools/profile_step.py > gpurun_out/profile_step_r1u.log 2>&1; grep -E "wall|total|3072" gpurun_out/profile_step_r1u.log | cut -c1-150
