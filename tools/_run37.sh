mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit_eval.py tests/test_gpu_widening.py -m gpu -q -x > gpurun_out/r2an_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2an_tests.log
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "topk" > gpurun_out/r2an_tests_cfg.log 2>&1; echo "cfg4 rc=$?"; tail -3 gpurun_out/r2an_tests_cfg.log
timeout 300 python tools/eval_shard_probe.py 2>&1 | tee gpurun_out/r2an_eval_shard_probe.txt
B="python bench.py --no-twin --no-sharded --no-cpu-baseline --steps 20 --warmup 5"
for R in 0 4 8 10; do
  MFB_CHUNK_RAMP=$R timeout 300 $B > gpurun_out/r2an_b_ramp$R.json 2>/dev/null
  python - <<PY
import json
b=json.load(open('gpurun_out/r2an_b_ramp$R.json')); print('ramp $R: us/step %.2f e2e %.1fM' % (b['ms_per_step']*1e3, b['e2e']['value']/1e6))
PY
done
