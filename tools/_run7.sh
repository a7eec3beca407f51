python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "cross or gated" 2>&1 | tail -2
python tools/run_cross_kernels.py
