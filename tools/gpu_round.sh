set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r1_gpu_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --steps 10 --warmup 3 --regime render --no-cpu-baseline > gpurun_out/bench_render.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --frames 200 --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:deform_wq -s 3 -c 1 -o gpurun_out/prof_deform python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --frames 200 --e2e-steps 1 > gpurun_out/ncu_deform.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hash_fwd -s 3 -c 1 -o gpurun_out/prof_hash python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --frames 200 --e2e-steps 1 > gpurun_out/ncu_hash.log 2>&1
ls -la gpurun_out
