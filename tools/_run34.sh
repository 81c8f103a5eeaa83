mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2ak_bench_final.json 2> gpurun_out/r2ak_bench_final.err; echo "bench default rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2ak_bench_reference_arm.json 2> gpurun_out/r2ak_bench_reference_arm.err; echo "ref arm rc=$?"
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-twin --repeats 3 --sharded-steps 32 --sharded-warmup 32"
timeout 600 $CMD > gpurun_out/r2ak_bench_for_ncu.json 2>/dev/null; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2ak_bench_launches.csv $CMD > gpurun_out/r2ak_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
EVAL_TC_ONLY=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tc_gemm|k_tc_rescore|k_tc_threshold" -s 4 -c 4 -o gpurun_out/r2ak_eval_tc python tools/eval_bench.py > gpurun_out/r2ak_ncu_eval.log 2>&1; echo "ncu eval rc=$?"
ls -la gpurun_out/r2ak_*
