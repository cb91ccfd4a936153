# quick GPU check: parity tests + one bench line (no profiler)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -60 > gpurun_out/gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench.err
cat gpurun_out/gpu_tests.log; python - <<'PY'
import json
for f in ("bench_train",):
    try:
        j=json.load(open("gpurun_out/%s.json"%f)); print(f, "%.1f M/s"%(j["value"]/1e6), j["kernels_ms"], "e2e %.1f M/s"%(j["e2e"]["value"]/1e6), {k:v for k,v in j.items() if k in ("network","render","train_step","cpu_baseline","ref_cuda")})
    except Exception as e: print(f, "failed", e)
PY
tail -5 gpurun_out/bench.err
