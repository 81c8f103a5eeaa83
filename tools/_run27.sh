mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit_eval.py -m gpu -q -x -k "tc or topk or tensor" > gpurun_out/r2ad_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2ad_tests.log
for V in libmfb200; do
  MFB_LIB_PATH=recommendation_gans_b200/lib/$V.so EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2ad_eval_$V.log 2>&1
  echo "=== $V rc=$? $(grep 'MFB_TC=1' gpurun_out/r2ad_eval_$V.log)"
done
MFB_LIB_PATH=recommendation_gans_b200/lib/var_timing.so EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2ad_eval_t.log 2>&1
grep "tc timing" gpurun_out/r2ad_eval_t.log | tail -4 | grep "cta 0" | cut -c1-350
V=libmfb200
MFB_LIB_PATH=recommendation_gans_b200/lib/$V.so EVAL_TC_ONLY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_tc_gemm" -c 24 --csv --log-file gpurun_out/r2ad_launches_$V.csv python tools/eval_bench.py > gpurun_out/r2ad_ncu_$V.log 2>&1
echo "== $V rc=$?"; python tools/ncu_summary.py gpurun_out/r2ad_launches_$V.csv | cut -c1-140
