mkdir -p gpurun_out
V=recommendation_gans_b200/lib/var_timing.so
for DBG in 0 1 17 33 49 64 65 81 113; do
  MFB_TC_DBG=$DBG MFB_LIB_PATH=$V EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2u_eval_dbg$DBG.log 2>&1
  echo "=== DBG=$DBG rc=$? $(grep 'MFB_TC=1' gpurun_out/r2u_eval_dbg$DBG.log)"
  grep "tc timing" gpurun_out/r2u_eval_dbg$DBG.log | tail -4
done
