mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_steps.py tests/test_gpu_fit_eval.py -m gpu -q -x -k "steps_match or many_chunks or fast_math or fit_predict" > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2s_tests.log
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -k "cfg1 or cfg2" > gpurun_out/r2s_tests_cfg.log 2>&1; echo "cfg1/2 rc=$?"; tail -4 gpurun_out/r2s_tests_cfg.log
for W in cfg1 cfg2; do for SM in 1 0; do
  MFB_SMALL_STEPS=$SM timeout 300 python bench.py --workload $W --no-twin --no-sharded --no-cpu-baseline --steps 300 --warmup 10 > gpurun_out/r2s_b_${W}_small$SM.json 2>gpurun_out/r2s_b_${W}_small$SM.err; echo "$W small=$SM rc=$?"
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s_b_*.json')):
    try:
        b=json.load(open(f)); print(f.split('r2s_b_')[1], 'us/step %.2f  %.1fM inter/s e2e %.1fM  kernels %s'%(b['ms_per_step']*1e3, b['value']/1e6, b['e2e']['value']/1e6, {k:round(v,1) for k,v in b['kernel_us_per_step'].items()}))
    except Exception as e: print(f,'ERR',e)
PY
