mkdir -p gpurun_out/m2
for i in 1 2 3; do
SOCCDPT_CONV_M2=1 timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/m2/on_$i.json 2> gpurun_out/m2/on_$i.err
SOCCDPT_CONV_M2=0 timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/m2/off_$i.json 2> gpurun_out/m2/off_$i.err
done
python - <<'PY'
import json
for i in (1,2,3):
  for n in ("on","off"):
    d=json.loads(open(f"gpurun_out/m2/{n}_{i}.json").read().strip().splitlines()[-1])
    print(n, i, round(d["value"]), round(d["ms_per_step"],3), 'e2e', round(d["e2e"]["ms_per_step"],3), d["kernels_ms_per_step"]["conv_tcgen05_kernel"], d["clocks"])
PY
