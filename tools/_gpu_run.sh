mkdir -p gpurun_out/dt
timeout 300 python -m pytest tests/test_gpu_ops.py -k "depth" -x -q -m gpu 2>&1 | tail -15 > gpurun_out/dt/pytest.log; cat gpurun_out/dt/pytest.log | tail -6
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -k "depth_tail_vs" -x -q -m gpu 2>&1 | tail -12 > gpurun_out/dt/memcheck.log; tail -4 gpurun_out/dt/memcheck.log
for v in default old kp64 ring8 ring5kp16; do
  if [ $v = default ]; then unset SOCCDPT_LIB; else export SOCCDPT_LIB=build/variants/$v/lib.so; fi
  echo "== $v"; timeout 200 python tools/bench_depth_tail.py 2>&1 | tail -2
done > gpurun_out/dt/ab.log 2>&1
cat gpurun_out/dt/ab.log
unset SOCCDPT_LIB
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/dt/bench.json 2> gpurun_out/dt/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/dt/bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["kernels_ms_per_step"])
PY
