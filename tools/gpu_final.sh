# final pass of a round: the whole GPU suite, the default bench line, the launch list of the bench command (kernel shares), one full capture of the dominant kernel
mkdir -p gpurun_out
T=${TAG:-r2_final}
date +%T
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread 2>&1 | tail -6 > gpurun_out/${T}_gpu_tests.log; cat gpurun_out/${T}_gpu_tests.log
date +%T
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
date +%T
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_bench.log 2>&1
date +%T
timeout 300 ncu --set full --import-source on --clock-control none -k "regex:deform_pf|transform_kernel" -s 6 -c 3 -o gpurun_out/${T}_headline python tools/prof_deform.py 1700 > gpurun_out/${T}_ncu_headline.log 2>&1; tail -2 gpurun_out/${T}_ncu_headline.log
date +%T
python - <<PY
import json
j=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("%.1f M/s"%(j["value"]/1e6), "ms/step %.3f"%j["ms_per_step"], j["kernels_ms"], "e2e %.1f M/s"%(j["e2e"]["value"]/1e6), "roof", j["roofline"]["frac"], "launches", j["gpu_launches"], j["clocks"])
for k in ("network","render","train_step","speedup_vs_ref_cuda"):
    v=j.get(k)
    if isinstance(v,dict): print(k, {a:b for a,b in v.items() if not isinstance(b,(list,dict,str))})
PY
