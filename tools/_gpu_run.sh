for v in default dt_r4_c3 dt_r5_c2 dt_r4_c2; do
  if [ $v = default ]; then unset SOCCDPT_LIB; else export SOCCDPT_LIB=build/variants/$v/lib.so; fi
  echo "== $v"; timeout 200 python tools/bench_depth_tail.py 2>&1 | tail -2
done
unset SOCCDPT_LIB
SOCCDPT_LIB=build/variants/dt_r4_c3/lib.so timeout 200 python -m pytest tests/test_gpu_ops.py -k depth -x -q -m gpu 2>&1 | tail -2
