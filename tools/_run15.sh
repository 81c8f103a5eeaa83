mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit_eval.py -m gpu -q > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2o_tests.log
for V in libmfb200 var_st4; do for CLU in 1 2; do
  MFB_LIB_PATH=recommendation_gans_b200/lib/$V.so MFB_TC_CLUSTER=$CLU EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2o_eval_${V}_$CLU.log 2>&1; echo "$V cluster=$CLU $(grep 'MFB_TC=1' gpurun_out/r2o_eval_${V}_$CLU.log)"
done; done
EVAL_TC_ONLY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2o_eval_launches.csv python tools/eval_bench.py > /dev/null 2>&1; python tools/ncu_summary.py gpurun_out/r2o_eval_launches.csv | grep "k_tc"
MFB_TC_CLUSTER=2 EVAL_TC_ONLY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_tc_gemm -c 8 --csv --log-file gpurun_out/r2o_eval_launches_cl2.csv python tools/eval_bench.py > /dev/null 2>&1; echo cluster2; python tools/ncu_summary.py gpurun_out/r2o_eval_launches_cl2.csv | grep "k_tc"
