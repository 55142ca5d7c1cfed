mkdir -p gpurun_out/alt
timeout 600 python -m pytest tests/test_gpu_conv_tcgen05.py tests/test_gpu_network.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/alt/pytest.log; tail -3 gpurun_out/alt/pytest.log
for i in 1 2; do
SOCCDPT_CONV_ALT=1 timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/alt/on_$i.json 2> gpurun_out/alt/on_$i.err
SOCCDPT_CONV_ALT=0 timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/alt/off_$i.json 2> gpurun_out/alt/off_$i.err
done
python - <<'PY'
import json
for i in (1,2):
  for n in ("on","off"):
    d=json.loads(open(f"gpurun_out/alt/{n}_{i}.json").read().strip().splitlines()[-1])
    print(n, i, round(d["value"]), round(d["ms_per_step"],3), d["kernels_ms_per_step"]["conv_tcgen05_kernel"], d["clocks"]["sm_mhz"])
PY
PYTHONPATH=. timeout 200 python tools/bench_conv.py 2>&1 | tail -40 > gpurun_out/alt/conv_on.log
SOCCDPT_CONV_ALT=0 PYTHONPATH=. timeout 200 python tools/bench_conv.py 2>&1 | tail -40 > gpurun_out/alt/conv_off.log
paste -d'|' gpurun_out/alt/conv_on.log gpurun_out/alt/conv_off.log | cut -c1-230 | head -45
