for v in 0 600 1100 1600; do
  if [ $v = 0 ]; then L=modaltune_b200/libmodaltune_b200.so; else L=build_exp/libmt_stag$v.so; fi
  echo "== stagger $v"
  MODALTUNE_B200_LIB=$L python tools/run_attn_kernels.py 10001 6 3 2
  MODALTUNE_B200_LIB=$L python tools/run_attn_kernels.py 32769 4 3 2
done
