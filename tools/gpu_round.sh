# profiling pass of a round: launch list of the default bench command + full captures of the top kernels (one GPU; never a bench number)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --frames 200 --e2e-steps 1 --train-images 32 --train-steps 3 --train-warmup 12 --render-frames 2"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deform_wq -s 3 -c 1 -o gpurun_out/prof_deform $B --no-render > gpurun_out/ncu_deform.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nerf_network_fwd -s 1 -c 1 -o gpurun_out/prof_netfwd $B > gpurun_out/ncu_netfwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:net_fwd_bwd -s 14 -c 1 -o gpurun_out/prof_netbwd $B > gpurun_out/ncu_netbwd.log 2>&1
ls -la gpurun_out
