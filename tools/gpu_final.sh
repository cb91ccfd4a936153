# final pass of a round: GPU parity tests, the default bench line, the comparison with the reference's own CUDA code (no profiler)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gpu_tests_v14.log
timeout 300 python bench.py > gpurun_out/bench_default_v14.json 2> gpurun_out/bench_v14.err
timeout 300 python tests/bench_ref_cuda.py > gpurun_out/vs_ref_cuda_v14.json 2> gpurun_out/vs_v14.err
cat gpurun_out/gpu_tests_v14.log; tail -n 2 gpurun_out/bench_v14.err; tail -n 2 gpurun_out/vs_v14.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/bench_default_v14.json")); print("%.1f M/s"%(j["value"]/1e6), j["kernels_ms"], "e2e %.1f"%(j["e2e"]["value"]/1e6), "render", j["render"]["ms_per_frame"], "train", j["train_step"]["ms_per_step"], "roof", j["roofline"]["frac"])
v=json.load(open("gpurun_out/vs_ref_cuda_v14.json")); print(v["speedup"])
PY
