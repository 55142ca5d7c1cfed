mkdir -p gpurun_out/r2J
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "window_attention_tma" > gpurun_out/r2J/t_attn.log 2>&1; echo "attn rc=$?"
tail -15 gpurun_out/r2J/t_attn.log
timeout 600 python -m pytest tests/test_gpu_conv_tcgen05.py tests/test_gpu_network.py -x -q -m gpu > gpurun_out/r2J/t_net.log 2>&1; echo "net rc=$?"
tail -5 gpurun_out/r2J/t_net.log
timeout 300 python bench.py > gpurun_out/r2J/bench_tma.json 2> gpurun_out/r2J/bench_tma.err; echo "bench rc=$?"
SOCCDPT_ATTN_TMA=0 timeout 300 python bench.py > gpurun_out/r2J/bench_old.json 2> gpurun_out/r2J/bench_old.err
python - <<'PY'
import json
for n in ("tma","old"):
    try:
        d=json.loads(open(f"gpurun_out/r2J/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["kernels_ms_per_step"])
    except Exception as e: print(n, "ERR", e)
PY
