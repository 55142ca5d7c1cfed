mkdir -p gpurun_out/fin6
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/fin6/pytest.log; tail -3 gpurun_out/fin6/pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/fin6/bench_tiny.json 2> gpurun_out/fin6/bench_tiny.err; echo "tiny rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/fin6/bench_tiny.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],3), 'e2e', round(d["e2e"]["value"]), d.get("model_frac_of_peak"), d["roofline"]["frac"], d["clocks"], d["kernels_ms_per_step"])
PY
