mkdir -p gpurun_out
for V in libmfb200 var_ss; do
  MFB_LIB_PATH=recommendation_gans_b200/lib/$V.so EVAL_TC_ONLY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_tc_" -c 60 --csv --log-file gpurun_out/r2z_launches_$V.csv python tools/eval_bench.py > gpurun_out/r2z_ncu_$V.log 2>&1
  echo "== $V rc=$?"; python tools/ncu_summary.py gpurun_out/r2z_launches_$V.csv | cut -c1-140
done
