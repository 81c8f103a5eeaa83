mkdir -p gpurun_out; rm -f gpurun_out/config_parity.jsonl
python -m pytest tests/test_gpu_configs.py tests/test_gpu_steps.py tests/test_gpu_fit_eval.py tests/test_gpu_rng.py -m gpu -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2c_tests.log
B="python bench.py --no-twin --no-sharded --no-cpu-baseline"
for K in 20 494; do
  $B --steps $K --warmup 5 > gpurun_out/r2c_b_ramp4_$K.json 2>/dev/null; echo "ramp4 $K rc=$?"
  MFB_CHUNK_RAMP=0 $B --steps $K --warmup 5 > gpurun_out/r2c_b_ramp0_$K.json 2>/dev/null
  MFB_CHUNK_RAMP=2 $B --steps $K --warmup 5 > gpurun_out/r2c_b_ramp2_$K.json 2>/dev/null
  MFB_CHUNK_RAMP=8 $B --steps $K --warmup 5 > gpurun_out/r2c_b_ramp8_$K.json 2>/dev/null
  MFB_LIB_PATH=recommendation_gans_b200/lib/var_spw1.so $B --steps $K --warmup 5 > gpurun_out/r2c_b_spw1_$K.json 2>/dev/null
  MFB_LIB_PATH=recommendation_gans_b200/lib/var_spw4.so $B --steps $K --warmup 5 > gpurun_out/r2c_b_spw4_$K.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c_b_*.json')):
    try:
        b=json.load(open(f))
        print(f.split('r2c_b_')[1], 'ms/step %.4f [%.4f..%.4f] e2e %.1fM upd %.1fus k_us %s'%(b['ms_per_step'], b['timing']['ms_per_step_min'], b['timing']['ms_per_step_max'], b['e2e']['value']/1e6, b['roofline']['us_per_launch'], {k:round(v,1) for k,v in b['kernel_us_per_step'].items()}))
    except Exception as e: print(f, 'ERR', e)
PY
