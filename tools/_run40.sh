mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2aq_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2aq_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2aq_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2aq_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2aq_bench_driver.json 2> gpurun_out/r2aq_bench_driver.err; echo "bench rc=$?"
timeout 900 python bench.py > gpurun_out/r2aq_bench_final.json 2> gpurun_out/r2aq_bench_final.err; echo "bench default rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2aq_bench_driver.json','gpurun_out/r2aq_bench_final.json'):
    b=json.load(open(f))
    print(f, 'value %.1fM e2e %.1fM ms/step %.4f frac %.3f launches %s' % (b['value']/1e6, b['e2e']['value']/1e6, b['ms_per_step'], b['roofline']['frac'], b.get('gpu_launches')))
    print('  eval', b['eval']['seconds'], b['eval']['roofline']['frac'], b['eval']['timing'].get('unkeyed_seconds'), 'zipf', b.get('zipf',{}).get('value'), 'sharded', b['sharded_train']['value'])
PY
