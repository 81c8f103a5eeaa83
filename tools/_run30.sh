mkdir -p gpurun_out
timeout 300 python tools/eval_shard_probe.py 2>&1 | tee gpurun_out/r2ag_eval_shard_probe.txt
timeout 900 python -m pytest tests/test_gpu_fit_eval.py tests/test_gpu_widening.py -m gpu -q -x > gpurun_out/r2ag_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ag_tests.log
