mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gpu_tests_v13.log
timeout 600 python bench.py > gpurun_out/bench_default_v13.json 2> gpurun_out/bench_v13.err
timeout 600 python tests/bench_ref_cuda.py > gpurun_out/vs_ref_cuda_v13.json 2> gpurun_out/vs_v13.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --frames 200 --e2e-steps 1 --train-images 32 --train-steps 3 --train-warmup 70 --render-frames 2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_v13.csv $B > gpurun_out/ncu_bench_v13.log 2>&1
cat gpurun_out/gpu_tests_v13.log; tail -n 3 gpurun_out/bench_v13.err; tail -n 3 gpurun_out/vs_v13.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/bench_default_v13.json")); print("%.1f M/s"%(j["value"]/1e6), j["kernels_ms"], "e2e %.1f"%(j["e2e"]["value"]/1e6), "render", j["render"]["ms_per_frame"], "train", j["train_step"]["ms_per_step"], "roof", j["roofline"]["frac"])
v=json.load(open("gpurun_out/vs_ref_cuda_v13.json")); print(v["speedup"])
PY
