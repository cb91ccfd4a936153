mkdir -p gpurun_out
T=${TAG:-r2}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_gpu_tests.log
cat gpurun_out/${T}_gpu_tests.log
timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
j=json.load(open("gpurun_out/${T}_bench.json"))
print("%.1f M/s"%(j["value"]/1e6), "ms/step %.3f"%j["ms_per_step"], j["kernels_ms"], "e2e %.1f M/s"%(j["e2e"]["value"]/1e6), "roof", j["roofline"]["frac"])
for k in ("network","render","train_step","cpu_baseline","ref_cuda"):
    v=j.get(k)
    if isinstance(v,dict): print(k, {a:b for a,b in v.items() if not isinstance(b,(list,dict,str))})
PY
