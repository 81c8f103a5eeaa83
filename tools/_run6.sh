mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_steps.py tests/test_gpu_sharded.py tests/test_gpu_rng.py -m gpu -q -x > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2f_tests.log
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -k "cfg3 or cfg1" > gpurun_out/r2f_tests_cfg.log 2>&1; echo "cfg tests rc=$?"; tail -3 gpurun_out/r2f_tests_cfg.log
B="python bench.py --no-twin --no-sharded --no-cpu-baseline"
for K in 20 494; do
 for IT in uniform zipf; do
  timeout 300 $B --items $IT --steps $K --warmup 5 > gpurun_out/r2f_b_il_${IT}_$K.json 2>/dev/null
  MFB_LIB_PATH=recommendation_gans_b200/lib/var_noil.so timeout 300 $B --items $IT --steps $K --warmup 5 > gpurun_out/r2f_b_noil_${IT}_$K.json 2>/dev/null
 done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_b_*.json')):
    try:
        b=json.load(open(f))
        print(f.split('r2f_b_')[1], 'ms/step %.4f [%.4f..%.4f] e2e %.1fM upd %.1fus k_us %s'%(b['ms_per_step'], b['timing']['ms_per_step_min'], b['timing']['ms_per_step_max'], b['e2e']['value']/1e6, b['roofline']['us_per_launch'], {k:round(v,1) for k,v in b['kernel_us_per_step'].items()}))
    except Exception as e: print(f, 'ERR', e)
PY
