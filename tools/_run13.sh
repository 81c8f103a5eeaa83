mkdir -p gpurun_out; rm -f gpurun_out/config_parity.jsonl
timeout 1800 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r02_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -22 gpurun_out/r02_gpu_tests.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_driver.json 2> gpurun_out/r02_bench_driver.err; echo "bench driver-config rc=$?"
timeout 1200 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench default rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-twin --repeats 3 --sharded-steps 32 --sharded-warmup 32 > /dev/null 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_update|k_forward" -s 300 -c 4 -f -o gpurun_out/r02_update_forward python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-sharded --no-twin --repeats 1 > /dev/null 2>&1; echo "ncu train rc=$?"
EVAL_TC_ONLY=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tc_gemm|k_tc_rescore|k_tc_threshold" -s 4 -c 4 -f -o gpurun_out/r02_eval_tc python tools/eval_bench.py > /dev/null 2>&1; echo "ncu eval rc=$?"
ls -la gpurun_out/*.ncu-rep
