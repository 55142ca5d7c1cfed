mkdir -p gpurun_out/r2W
timeout 600 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/r2W/bench_overlap.json 2> gpurun_out/r2W/bench_overlap.err; echo "bench rc=$?"
SOCCDPT_STREAM_OVERLAP=0 timeout 300 python bench.py > gpurun_out/r2W/bench_nooverlap.json 2> gpurun_out/r2W/bench_nooverlap.err
python - <<'PY'
import json
for n in ("overlap","nooverlap"):
    d=json.loads(open(f"gpurun_out/r2W/bench_{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"])
PY
