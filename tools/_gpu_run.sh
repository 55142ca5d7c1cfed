mkdir -p gpurun_out/mg2
for n in 8 4 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --gather --no-cpu-baseline > gpurun_out/mg2/bench_tiny_${n}gpu.json 2> gpurun_out/mg2/bench_tiny_${n}gpu.err; echo "n=$n rc=$?"
done
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/mg2/bench_tiny_1gpu.json 2> gpurun_out/mg2/bench_tiny_1gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --model hybrid_384 --no-cpu-baseline > gpurun_out/mg2/bench_hybrid_8gpu.json 2> gpurun_out/mg2/bench_hybrid_8gpu.err; echo "hybrid rc=$?"
python - <<'PY'
import json
for f in ("bench_tiny_1gpu","bench_tiny_2gpu","bench_tiny_4gpu","bench_tiny_8gpu","bench_hybrid_8gpu"):
    try:
        d=json.loads(open(f"gpurun_out/mg2/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), 'e2e', round(d["e2e"]["value"]) if d.get("e2e") else None, d.get("nccl_gather_masks_ms"), d["clocks"])
    except Exception as e: print(f, 'ERR', e)
PY
