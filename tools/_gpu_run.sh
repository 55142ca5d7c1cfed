mkdir -p gpurun_out/cap
timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv_tcgen05 --launch-skip 52 -c 1 -o gpurun_out/cap/s0_qkv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/cap/ncu_a.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv_tcgen05 --launch-skip 101 -c 1 -o gpurun_out/cap/depth_conv0 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/cap/ncu_b.log 2>&1
ls -la gpurun_out/cap
