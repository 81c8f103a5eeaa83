mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit_eval.py tests/test_gpu_widening.py -m gpu -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2j_tests.log
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -k "topk_on_trained" > gpurun_out/r2j_tests_cfg.log 2>&1; echo "cfg4 rc=$?"; tail -3 gpurun_out/r2j_tests_cfg.log
EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2j_eval.log 2>&1; grep "MFB_TC=1\|candidate" gpurun_out/r2j_eval.log
EVAL_SKEW=2 EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2j_eval_skew.log 2>&1; grep "MFB_TC=1\|candidate" gpurun_out/r2j_eval_skew.log
python bench.py --no-sharded --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2j_bench.json 2>/dev/null
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2j_bench.json'))
for name,d in (('uniform',b),('zipf',b.get('zipf',{}))):
    e=d['eval']; print(name,'train ms/step %.4f | eval %.3f ms frac %.3f first %.3f unkeyed %.3f e2e %.1f ms'%(d['ms_per_step'], e['seconds']*1e3, e['roofline']['frac'], e['timing']['first_call_seconds']*1e3, e['timing']['unkeyed_seconds']*1e3, e['e2e']['seconds']*1e3), e['kernel'][-40:])
PY
