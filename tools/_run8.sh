mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fit_eval.py -m gpu -q -k "keyed or tc_topk or fit_predict" > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2h_tests.log
B="python bench.py --no-twin --no-sharded --no-cpu-baseline"
for V in libmfb200 var_minb12 var_minb14 var_minb8 var_w2; do
  MFB_LIB_PATH=recommendation_gans_b200/lib/$V.so timeout 300 $B --steps 494 --warmup 5 > gpurun_out/r2h_b_${V}.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2h_b_*.json')):
    try:
        b=json.load(open(f))
        print(f.split('r2h_b_')[1], 'ms/step %.4f upd %.1fus fwd %.1f | eval %.3f ms (first %.3f unkeyed %.3f) e2e %.1f ms'%(b['ms_per_step'], b['roofline']['us_per_launch'], b['kernel_us_per_step']['forward'], b['eval']['seconds']*1e3, b['eval']['timing']['first_call_seconds']*1e3, b['eval']['timing']['unkeyed_seconds']*1e3, b['eval']['e2e']['seconds']*1e3))
    except Exception as e: print(f, 'ERR', e)
PY
