mkdir -p gpurun_out/bt
timeout 300 python -m pytest tests/test_gpu_block_tail.py tests/test_gpu_network.py -x -q -m gpu 2>&1 | tail -2
echo "== new"; S2=1 timeout 200 python tools/bench_block_tail.py 2>&1 | tail -6
echo "== prev"; S2=1 SOCCDPT_LIB=build/variants/prev/lib.so timeout 200 python tools/bench_block_tail.py 2>&1 | tail -6
SOCCDPT_LIB=build/variants/trace/lib.so timeout 300 python tools/trace_block_tail.py > gpurun_out/bt/trace3.log 2>&1
grep -A3 "^== S0 mlp\|^== S0 proj" gpurun_out/bt/trace3.log | cut -c1-300
for i in 1 2; do
timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/bt/new_$i.json 2> gpurun_out/bt/new_$i.err
SOCCDPT_LIB=build/variants/prev/lib.so timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/bt/prev_$i.json 2> gpurun_out/bt/prev_$i.err
done
python - <<'PY'
import json
for i in (1,2):
  for n in ("new","prev"):
    d=json.loads(open(f"gpurun_out/bt/{n}_{i}.json").read().strip().splitlines()[-1])
    print(n, i, round(d["value"]), round(d["ms_per_step"],3), d["kernels_ms_per_step"]["swin_block_tail_kernel"], d["clocks"]["sm_mhz"])
PY
