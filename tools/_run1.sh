python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1p.log 2>&1; tail -2 gpurun_out/pytest_r1p.log
python bench.py > gpurun_out/bench_r1p.log 2>&1; tail -1 gpurun_out/bench_r1p.log | cut -c1-300
for v in 0 1 2 3; do
  if [ $v = 0 ]; then L=modaltune_b200/libmodaltune_b200.so; else L=build_exp/libmt_poly$v.so; fi
  echo "== poly $v" >> gpurun_out/poly_exp.log
  MODALTUNE_B200_LIB=$L python tools/run_attn_kernels.py 10001 6 3 2 >> gpurun_out/poly_exp.log 2>&1
  MODALTUNE_B200_LIB=$L python tools/run_attn_kernels.py 32769 4 3 2 >> gpurun_out/poly_exp.log 2>&1
  [ $v != 0 ] && MODALTUNE_B200_LIB=$L python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "tcgen05" 2>&1 | tail -3 >> gpurun_out/poly_exp.log
done
cat gpurun_out/poly_exp.log
