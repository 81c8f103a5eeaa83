mkdir -p gpurun_out
V=recommendation_gans_b200/lib/var_timing.so
for SC in 0 0.0078125 1.0; do
  MFB_TC_DBG=17 MFB_LIB_PATH=$V EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py $SC > gpurun_out/r2x_eval_sc$SC.log 2>&1
  echo "=== scale=$SC rc=$?"
  grep "tc timing COLLECT cta 0" gpurun_out/r2x_eval_sc$SC.log | tail -1
done
