# round-2 check: all GPU tests + the default bench line (tag = $TAG)
mkdir -p gpurun_out
T=${TAG:-r2}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${T}_gpu_tests.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_gpu_tests.log; tail -n 3 gpurun_out/${T}_bench.err
python - <<PY
import json
j=json.load(open("gpurun_out/${T}_bench.json"))
print("%.1f M/s"%(j["value"]/1e6), "ms/step", j["ms_per_step"], j["kernels_ms"], "e2e %.1f M/s"%(j["e2e"]["value"]/1e6), "roof", j["roofline"]["frac"])
for k in ("network","render","train_step","ref_cuda","speedup_vs_ref_cuda","stress","ingest"):
    print(k, json.dumps(j.get(k))[:900])
PY
