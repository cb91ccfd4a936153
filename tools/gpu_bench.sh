# one bench line (no profiler) + optional ncu capture of the kernels named in NCU_K (regex) from tools/bwd_ab.py
mkdir -p gpurun_out
T=${TAG:-r2_bench}
date +%T
if [ -n "$NCU_K" ]; then
timeout 300 ncu --set full --import-source on --clock-control none -k regex:$NCU_K -s ${NCU_SKIP:-2} -c 1 -o gpurun_out/${T}_ncu python ${NCU_CMD:-tools/bwd_ab.py} > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log
date +%T
fi
timeout 1200 python bench.py ${BENCH_ARGS} > gpurun_out/${T}.json 2> gpurun_out/${T}.err
date +%T
tail -3 gpurun_out/${T}.err
python - <<PY
import json
j=json.loads(open("gpurun_out/${T}.json").read().strip().splitlines()[-1])
print("%.1f M/s"%(j["value"]/1e6), "ms/step %.3f"%j["ms_per_step"], j["kernels_ms"], "e2e %.1f M/s"%(j["e2e"]["value"]/1e6), "roof", j["roofline"]["frac"])
for k in ("network","render","train_step","ref_cuda","speedup_vs_ref_cuda","e2e_inference","stress"):
    v=j.get(k)
    if isinstance(v,dict): print(k, {a:b for a,b in v.items() if not isinstance(b,(list,dict,str))})
print("train ref", j.get("train_step",{}).get("ref_cuda",{}).get("ms_per_step"), "render ref", j.get("render",{}).get("ref_cuda",{}).get("ms_per_frame"))
PY
