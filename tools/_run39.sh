mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit_eval.py tests/test_gpu_widening.py -m gpu -q -x > gpurun_out/r2ap_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2ap_tests.log
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "topk" > gpurun_out/r2ap_tests_cfg.log 2>&1; echo "cfg4 rc=$?"; tail -3 gpurun_out/r2ap_tests_cfg.log
timeout 300 python tools/eval_shard_probe.py 2>&1 | tee gpurun_out/r2ap_eval_shard_probe.txt
MFB_TC_FUSED_THR=0 timeout 300 python tools/eval_shard_probe.py 2>&1 | tee gpurun_out/r2ap_eval_shard_probe_unfused.txt
EVAL_TC_ONLY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_tc_|k_topk" -c 40 --csv --log-file gpurun_out/r2ap_launches.csv python tools/eval_bench.py > gpurun_out/r2ap_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2ap_launches.csv | cut -c1-140
