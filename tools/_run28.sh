mkdir -p gpurun_out
for V in libmfb200 var_spin var_st3; do
  MFB_LIB_PATH=recommendation_gans_b200/lib/$V.so EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2ae_eval_$V.log 2>&1
  echo "=== $V rc=$? $(grep 'MFB_TC=1' gpurun_out/r2ae_eval_$V.log)"
done
EVAL_SKEW=1 EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2ae_eval_skew.log 2>&1; echo "skew $(grep 'MFB_TC=1\|candidate' gpurun_out/r2ae_eval_skew.log)"
timeout 1500 python -m pytest tests/test_gpu_fit_eval.py tests/test_gpu_widening.py tests/test_gpu_configs.py tests/test_gpu_entry_point.py -m gpu -q -x > gpurun_out/r2ae_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ae_tests.log
