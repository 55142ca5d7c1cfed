mkdir -p gpurun_out/mg
for n in 8 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --gather > gpurun_out/mg/bench_tiny_${n}gpu.json 2> gpurun_out/mg/bench_tiny_${n}gpu.err; echo "n=$n rc=$?"
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --model base_384 --stream-frames 4096 > gpurun_out/mg/stream_base384_8gpu.json 2> gpurun_out/mg/stream_base384_8gpu.err; echo "stream rc=$?"
python - <<'PY'
import json
for f in ("bench_tiny_8gpu","bench_tiny_2gpu","stream_base384_8gpu"):
    try:
        d=json.loads(open(f"gpurun_out/mg/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), 'e2e', round(d["e2e"]["value"]) if d.get("e2e") else None, d.get("gather"), d["clocks"])
    except Exception as e: print(f, 'ERR', e)
PY
