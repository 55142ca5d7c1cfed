mkdir -p gpurun_out/fin3
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/fin3/pytest.log; tail -3 gpurun_out/fin3/pytest.log
K='conv_tcgen05|window_attention|swin_block_tail|layernorm|patch_embed|upsample|depth_tail|seg_finish|unproject|grid_expand|resize_tables|ln_res'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"$K" -s 97 -c 110 --csv --log-file gpurun_out/fin3/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/fin3/ncu1.log 2>&1
timeout 400 python bench.py > gpurun_out/fin3/bench_tiny.json 2> gpurun_out/fin3/bench_tiny.err; echo "tiny rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/fin3/bench_tiny.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],3), 'e2e', round(d["e2e"]["value"]), d.get("model_frac_of_peak"), d["roofline"]["frac"], d["clocks"], d["kernels_ms_per_step"])
PY
