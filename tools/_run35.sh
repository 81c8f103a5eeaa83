mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_rng.py -m gpu -q -x > gpurun_out/r2al_tests_2gpu.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2al_tests_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2al_bench_2gpu.json 2> gpurun_out/r2al_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2al_bench_2gpu.json'))
print('N=2 value %.1fM ms/step %.4f' % (b['value']/1e6, b['ms_per_step']))
print('eval', b['eval']['value'], b['eval']['seconds'], b['eval'].get('rank_seconds'))
print('sharded', b['sharded_train']['value'], b['sharded_train']['ms_per_step'])
PY
