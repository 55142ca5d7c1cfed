mkdir -p gpurun_out/fin4
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/fin4/pytest.log; tail -3 gpurun_out/fin4/pytest.log
K='conv_tcgen05|window_attention|swin_block_tail|layernorm|patch_embed|upsample|depth_tail|seg_finish|unproject|grid_expand|resize_tables|ln_res'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"$K" -s 97 -c 110 --csv --log-file gpurun_out/fin4/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/fin4/ncu1.log 2>&1
SOCCDPT_CONV_NSPLIT=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -s 97 -c 110 --csv --log-file gpurun_out/fin4/launches_nsplit0.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/fin4/ncu2.log 2>&1
timeout 400 python bench.py > gpurun_out/fin4/bench_tiny.json 2> gpurun_out/fin4/bench_tiny.err; echo "tiny rc=$?"
timeout 400 python bench.py --model base_384 --no-cpu-baseline > gpurun_out/fin4/bench_base_384.json 2> gpurun_out/fin4/bench_base_384.err; echo "base rc=$?"
timeout 400 python bench.py --model hybrid_384 --no-cpu-baseline > gpurun_out/fin4/bench_hybrid_384.json 2> gpurun_out/fin4/bench_hybrid_384.err; echo "hybrid rc=$?"
PYTHONPATH=. timeout 200 python tools/bench_models.py --version 1 --model dpt_swin2_tiny_256 --batch 64 > gpurun_out/fin4/v1.log 2>&1; grep "frames/s" gpurun_out/fin4/v1.log
PYTHONPATH=. timeout 200 python tools/bench_latency.py > gpurun_out/fin4/latency.log 2>&1; tail -4 gpurun_out/fin4/latency.log
python - <<'PY'
import json
for n in ("tiny","base_384","hybrid_384"):
    d=json.loads(open(f"gpurun_out/fin4/bench_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],3), 'e2e', round(d["e2e"]["value"]), d.get("model_frac_of_peak"), d["roofline"]["frac"], d["clocks"], d["kernels_ms_per_step"])
PY
