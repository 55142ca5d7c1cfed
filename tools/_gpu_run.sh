mkdir -p gpurun_out/r2V
timeout 200 python tools/bench_conv.py 2>&1 | tee gpurun_out/r2V/bench_conv.log
ONLY="S2 fc1" timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tcgen05 -s 5 -c 1 -o gpurun_out/r2V/s2fc1 python tools/bench_conv.py > gpurun_out/r2V/ncu.log 2>&1; echo "ncu rc=$?"
