python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "graphed" 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1v.log 2>&1; tail -1 gpurun_out/bench_r1v.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
