mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2ar_bench_8gpu.json 2> gpurun_out/r2ar_bench_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2ar_bench_8gpu.json'))
print('N=8 value %.1fM ms/step %.4f' % (b['value']/1e6, b['ms_per_step']))
print('eval', b['eval']['value'], b['eval']['seconds'], b['eval']['timing'].get('rank_seconds'))
print('sharded', b['sharded_train']['value'], b['sharded_train']['ms_per_step'])
PY
tail -3 gpurun_out/r2ar_bench_8gpu.err
