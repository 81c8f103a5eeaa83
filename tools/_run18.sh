mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_steps.py -m gpu -q -x > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2r_tests.log
B="python bench.py --no-twin --no-sharded --no-cpu-baseline"
for IT in uniform zipf; do for K in 20 494; do
  timeout 300 $B --items $IT --steps $K --warmup 5 > gpurun_out/r2r_b_${IT}_$K.json 2>/dev/null
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2r_b_*.json')):
    b=json.load(open(f)); print(f.split('r2r_b_')[1], 'ms/step %.4f upd %.1fus fwd %.1f frac %.3f'%(b['ms_per_step'], b['roofline']['us_per_launch'], b['kernel_us_per_step']['forward'], b['roofline']['frac']))
PY
