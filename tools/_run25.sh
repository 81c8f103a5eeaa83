mkdir -p gpurun_out
V=recommendation_gans_b200/lib/var_timing.so
MFB_LIB_PATH=$V EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2aa_eval.log 2>&1
echo "=== rc=$? $(grep 'MFB_TC=1' gpurun_out/r2aa_eval.log)"
grep "tc timing" gpurun_out/r2aa_eval.log | tail -4 | cut -c1-260
