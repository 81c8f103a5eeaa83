mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2aj_tests.log 2>&1; echo "tests rc=$?"; tail -14 gpurun_out/r2aj_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2aj_bench_driver.json 2> gpurun_out/r2aj_bench_driver.err; echo "bench rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2aj_bench_driver.json'))
print('value %.1fM e2e %.1fM ms/step %.4f roofline %s' % (b['value']/1e6, b['e2e']['value']/1e6, b['ms_per_step'], b['roofline']))
print('eval', {k: b['eval'][k] for k in b['eval'] if k in ('value','seconds','roofline','e2e','first_call','unkeyed')})
print('zipf', b.get('zipf', {}).get('value'), 'sharded', b.get('sharded_train', {}).get('value'), b.get('sharded_train', {}).get('ms_per_step'))
print('cpu', b.get('cpu_baseline'))
PY
