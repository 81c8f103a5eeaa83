mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit_eval.py tests/test_gpu_widening.py -m gpu -q -x > gpurun_out/r2as_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2as_tests.log
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "topk" > gpurun_out/r2as_tests_cfg.log 2>&1; echo "cfg4 rc=$?"; tail -3 gpurun_out/r2as_tests_cfg.log
timeout 300 python tools/eval_shard_probe.py 2>&1 | tee gpurun_out/r2as_eval_shard_probe.txt
