# quick GPU check: parity tests + one bench line per regime (no profiler)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --regime render --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_render.json 2>> gpurun_out/bench.err
cat gpurun_out/gpu_tests.log; python - <<'PY'
import json
for f in ("bench_train","bench_render"):
    try:
        j=json.load(open("gpurun_out/%s.json"%f)); print(f, "%.1f M/s"%(j["value"]/1e6), j["kernels_ms"], "e2e %.1f M/s"%(j["e2e"]["value"]/1e6), {k:v for k,v in j.items() if k in ("network","render","cpu_baseline","ref_cuda")})
    except Exception as e: print(f, "failed", e)
PY
tail -5 gpurun_out/bench.err
