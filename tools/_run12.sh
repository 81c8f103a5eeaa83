mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit_eval.py tests/test_gpu_widening.py -m gpu -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2l_tests.log
EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2l_eval.log 2>&1; grep "MFB_TC=1\|candidate" gpurun_out/r2l_eval.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2l_eval_launches.csv env EVAL_TC_ONLY=1 python tools/eval_bench.py > /dev/null 2>&1; python tools/ncu_summary.py gpurun_out/r2l_eval_launches.csv | grep -v "array\|Generator\|arange"
