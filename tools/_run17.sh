for ST in 2 3 4 6 8; do
  MFB_TC_SAMPLE_STEP=$ST EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2q_eval_$ST.log 2>&1; echo "sample_step=$ST $(grep 'MFB_TC=1\|candidate' gpurun_out/r2q_eval_$ST.log | tr '\n' ' ')"
done
