# round 2, backward side: timing of the scatter / network backward kernels + their parity tests
mkdir -p gpurun_out
T=${TAG:-r2_bwd}
date +%T > gpurun_out/${T}.log
timeout 300 python tools/bwd_ab.py ${BWD_ARGS} >> gpurun_out/${T}.log 2>&1
date +%T >> gpurun_out/${T}.log
timeout 900 python -m pytest ${TESTS:-tests/test_gpu_hash.py tests/test_gpu_network.py tests/test_gpu_network_ref.py tests/test_gpu_trainer.py} -m gpu -x -q --durations=6 2>&1 | tail -25 >> gpurun_out/${T}.log
date +%T >> gpurun_out/${T}.log
if [ -n "$NCU" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/bwd_ab.py > /dev/null 2>&1
grep -E "net_fwd_bwd|net_wgrad|hash_bwd|wgrad_finish" gpurun_out/${T}_launches.csv | awk -F'","' '{print $5, $NF}' | sed 's/(.*)//' | tail -4 >> gpurun_out/${T}.log
date +%T >> gpurun_out/${T}.log
fi
cat gpurun_out/${T}.log
