mkdir -p gpurun_out
for V in libmfb200 var_ts_sleep var_ss_elect var_ss_elect_sleep var_ss; do
  MFB_LIB_PATH=recommendation_gans_b200/lib/$V.so EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2ab_eval_$V.log 2>&1
  echo "=== $V rc=$? $(grep 'MFB_TC=1' gpurun_out/r2ab_eval_$V.log)"
done
MFB_LIB_PATH=recommendation_gans_b200/lib/var_ss_elect_timing.so EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2ab_eval_sst.log 2>&1
grep "tc timing" gpurun_out/r2ab_eval_sst.log | tail -4 | grep "cta 0" | cut -c1-330
for V in var_ss_elect var_ts_sleep; do
  MFB_LIB_PATH=recommendation_gans_b200/lib/$V.so EVAL_TC_ONLY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_tc_gemm" -c 24 --csv --log-file gpurun_out/r2ab_launches_$V.csv python tools/eval_bench.py > gpurun_out/r2ab_ncu_$V.log 2>&1
  echo "== $V rc=$?"; python tools/ncu_summary.py gpurun_out/r2ab_launches_$V.csv | cut -c1-140
done
