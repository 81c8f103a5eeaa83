mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit_eval.py -m gpu -q -x -k "tc or topk or tensor" > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2w_tests.log
EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2w_eval.log 2>&1; echo "rc=$? $(grep 'MFB_TC=1\|candidate' gpurun_out/r2w_eval.log)"
MFB_LIB_PATH=recommendation_gans_b200/lib/var_ss.so EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2w_eval_ss.log 2>&1; echo "SS rc=$? $(grep 'MFB_TC=1' gpurun_out/r2w_eval_ss.log)"
V=recommendation_gans_b200/lib/var_timing.so
for DBG in 0 17; do
  MFB_TC_DBG=$DBG MFB_LIB_PATH=$V EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2w_eval_dbg$DBG.log 2>&1
  echo "=== DBG=$DBG rc=$? $(grep 'MFB_TC=1' gpurun_out/r2w_eval_dbg$DBG.log)"
  grep "tc timing" gpurun_out/r2w_eval_dbg$DBG.log | tail -4
done
