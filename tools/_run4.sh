mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rng.py tests/test_gpu_steps.py -m gpu -q -x > gpurun_out/r2d_tests_a.log 2>&1; echo "rng+steps rc=$?"; tail -3 gpurun_out/r2d_tests_a.log
timeout 600 python -m pytest tests/test_gpu_fit_eval.py -m gpu -q > gpurun_out/r2d_tests_b.log 2>&1; echo "fit_eval rc=$?"; tail -6 gpurun_out/r2d_tests_b.log
for CLU in 2 1; do
  MFB_TC_CLUSTER=$CLU EVAL_TC_ONLY=1 timeout 300 python tools/eval_bench.py > gpurun_out/r2d_eval_cl$CLU.log 2>&1; echo "eval cluster=$CLU rc=$?"; grep "MFB_TC=1" gpurun_out/r2d_eval_cl$CLU.log
done
B="python bench.py --no-twin --no-sharded --no-cpu-baseline"
for K in 20 494; do
  for R in 0 4 8 16; do
    MFB_CHUNK_RAMP=$R timeout 300 $B --steps $K --warmup 5 > gpurun_out/r2d_b_ramp${R}_$K.json 2>/dev/null
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_b_*.json')):
    try:
        b=json.load(open(f))
        print(f.split('r2d_b_')[1], 'ms/step %.4f [%.4f..%.4f] e2e %.1fM upd %.1fus eval %.3fms k_us %s'%(b['ms_per_step'], b['timing']['ms_per_step_min'], b['timing']['ms_per_step_max'], b['e2e']['value']/1e6, b['roofline']['us_per_launch'], b['eval']['seconds']*1e3, {k:round(v,1) for k,v in b['kernel_us_per_step'].items()}))
    except Exception as e: print(f, 'ERR', e)
PY
