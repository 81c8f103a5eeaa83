mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rng.py tests/test_gpu_steps.py tests/test_gpu_fit_eval.py -m gpu -q -x > gpurun_out/r2ai_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2ai_tests.log
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "cfg3 or cfg5" > gpurun_out/r2ai_tests_cfg.log 2>&1; echo "cfg rc=$?"; tail -4 gpurun_out/r2ai_tests_cfg.log
B="python bench.py --no-twin --no-sharded --no-cpu-baseline"
for K in 20; do
  timeout 300 $B --steps $K --warmup 5 > gpurun_out/r2ai_b_$K.json 2>gpurun_out/r2ai_b_$K.err; echo "bench $K rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2ai_b_*.json')):
    try:
        b=json.load(open(f)); print(f, 'us/step %.2f  %.1fM inter/s e2e %.1fM  kernels %s'%(b['ms_per_step']*1e3, b['value']/1e6, b['e2e']['value']/1e6, {k:round(v,1) for k,v in b['kernel_us_per_step'].items()}))
    except Exception as e: print(f,'ERR',e)
PY
