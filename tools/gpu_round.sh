mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gpu_tests_v12.log
timeout 600 python bench.py > gpurun_out/bench_default_v12.json 2> gpurun_out/bench_v12.err
timeout 600 python tests/bench_ref_cuda.py > gpurun_out/vs_ref_cuda_v12.json 2> gpurun_out/vs_v12.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --frames 200 --e2e-steps 1 --train-images 32 --train-steps 3 --train-warmup 70 --render-frames 2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_v12.csv $B > gpurun_out/ncu_bench_v12.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deform_wq -s 3 -c 1 -o gpurun_out/prof_deform_v12 $B --no-render > gpurun_out/ncu_deform_v12.log 2>&1
cat gpurun_out/gpu_tests_v12.log; tail -n 3 gpurun_out/bench_v12.err; tail -n 3 gpurun_out/vs_v12.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/bench_default_v12.json")); print("%.1f M/s"%(j["value"]/1e6), j["kernels_ms"], "e2e %.1f"%(j["e2e"]["value"]/1e6), "render", j["render"]["ms_per_frame"], "train", j["train_step"]["ms_per_step"], "roof", j["roofline"]["frac"])
v=json.load(open("gpurun_out/vs_ref_cuda_v12.json")); print(v["speedup"])
PY
