mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2am_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2am_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2am_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2am_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2am_bench_driver.json 2> gpurun_out/r2am_bench_driver.err; echo "bench rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2am_bench_driver.json'))
print('value %.1fM e2e %.1fM ms/step %.4f frac %.3f launches %s' % (b['value']/1e6, b['e2e']['value']/1e6, b['ms_per_step'], b['roofline']['frac'], b.get('gpu_launches')))
print('eval', b['eval']['seconds'], b['eval']['roofline']['frac'])
print({k:round(v,1) for k,v in b['kernel_us_per_step'].items()})
PY
