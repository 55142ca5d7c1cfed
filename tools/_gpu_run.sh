mkdir -p gpurun_out/pre
timeout 600 python -m pytest tests/test_gpu_conv_tcgen05.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/pre/new_$i.json 2> gpurun_out/pre/new_$i.err
SOCCDPT_LIB=build/variants/prev/lib.so timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/pre/prev_$i.json 2> gpurun_out/pre/prev_$i.err
done
python - <<'PY'
import json
for i in (1,2):
  for n in ("new","prev"):
    d=json.loads(open(f"gpurun_out/pre/{n}_{i}.json").read().strip().splitlines()[-1])
    print(n, i, round(d["value"]), round(d["ms_per_step"],3), d["kernels_ms_per_step"]["conv_tcgen05_kernel"], d["clocks"]["sm_mhz"])
PY
