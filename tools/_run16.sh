mkdir -p gpurun_out
MFB_TC_CLUSTER=2 timeout 300 python -m pytest tests/test_gpu_fit_eval.py -m gpu -q -x -k "tc_ or keyed" > gpurun_out/r2p_tests.log 2>&1; echo "pair tests rc=$?"; tail -12 gpurun_out/r2p_tests.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader
for CLU in 1 2; do
  MFB_TC_CLUSTER=$CLU EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2p_eval_$CLU.log 2>&1; echo "cluster=$CLU rc=$? $(grep 'MFB_TC=1' gpurun_out/r2p_eval_$CLU.log)"
done
MFB_TC_CLUSTER=2 EVAL_TC_ONLY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_tc_gemm -c 8 --csv --log-file gpurun_out/r2p_launches_cl2.csv python tools/eval_bench.py > /dev/null 2>&1; python tools/ncu_summary.py gpurun_out/r2p_launches_cl2.csv | grep "k_tc"
