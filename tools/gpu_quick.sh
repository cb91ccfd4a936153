# quick GPU check: parity tests + one bench line per regime (no profiler)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gpu_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train.json 2> gpurun_out/bench.err
python bench.py --steps 10 --warmup 3 --regime render --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_render.json 2>> gpurun_out/bench.err
cat gpurun_out/gpu_tests.log; python - <<'PY'
import json
for f in ("bench_train","bench_render"):
    try:
        j=json.load(open("gpurun_out/%s.json"%f)); print(f, "%.1f M/s"%(j["value"]/1e6), j["kernels_ms"], "e2e %.1f M/s"%(j["e2e"]["value"]/1e6))
    except Exception as e: print(f, "failed", e)
PY
tail -5 gpurun_out/bench.err
