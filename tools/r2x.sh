mkdir -p gpurun_out
T=${TAG:-r2c}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests/test_gpu_deform.py tests/test_gpu_vs_ref_cuda.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
timeout 600 python tools/tune_deform.py ${VARIANTS:-2,20,d} d 1700 > gpurun_out/${T}_tune.log 2>&1
cat gpurun_out/${T}_tune.log
[ -n "$NONCU" ] || timeout 600 ncu --set full --import-source on --clock-control none -k "regex:deform_pf|transform_kernel" -s 4 -c 2 -o gpurun_out/${T}_deform_pf python tools/prof_deform.py ${PROF_FRAMES:-1700} > gpurun_out/${T}_ncu.log 2>&1
[ -n "$NONCU" ] || tail -5 gpurun_out/${T}_ncu.log
if [ -n "$OVERLAP" ]; then timeout 600 python tools/overlap_exp.py 1700 > gpurun_out/${T}_overlap.log 2>&1; cat gpurun_out/${T}_overlap.log; fi
