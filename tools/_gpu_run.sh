mkdir -p gpurun_out/sw
timeout 600 python -m pytest tests/test_gpu_conv_tcgen05.py tests/test_gpu_ops.py tests/test_gpu_network.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/sw/pytest.log; tail -4 gpurun_out/sw/pytest.log
SOCCDPT_LIB=build/variants/trace/lib.so timeout 300 python tools/trace_conv.py > gpurun_out/sw/trace.log 2>&1
grep -A5 "^==" gpurun_out/sw/trace.log | cut -c1-520
timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/sw/bench.json 2> gpurun_out/sw/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/sw/bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],3), d["kernels_ms_per_step"], d["clocks"]["sm_mhz"])
PY
