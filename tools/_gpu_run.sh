mkdir -p gpurun_out/r2T
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2T/t_all.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/r2T/t_all.log
timeout 300 python bench.py > gpurun_out/r2T/bench.json 2> gpurun_out/r2T/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2T/bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["kernels_ms_per_step"], d["roofline"]["frac"], d["model_frac_of_peak"], d["e2e"]["value"])
PY
timeout 200 python tools/bench_block_tail.py 2>&1 | tail -8
