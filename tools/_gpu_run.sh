mkdir -p gpurun_out/up
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/up/pytest.log; tail -4 gpurun_out/up/pytest.log
for i in 1 2; do
SOCCDPT_FOLD_UPSAMPLE=1 timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/up/on_$i.json 2> gpurun_out/up/on_$i.err
SOCCDPT_FOLD_UPSAMPLE=0 timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/up/off_$i.json 2> gpurun_out/up/off_$i.err
done
python - <<'PY'
import json
for i in (1,2):
  for n in ("on","off"):
    d=json.loads(open(f"gpurun_out/up/{n}_{i}.json").read().strip().splitlines()[-1])
    print(n, i, round(d["value"]), round(d["ms_per_step"],3), d["kernels_ms_per_step"]["conv_tcgen05_kernel"], d["kernels_ms_per_step"]["upsample"], d["clocks"]["sm_mhz"], d["gpu_launches"])
PY
