mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -k "nccl_two_gpus" > gpurun_out/r2k_tests_2gpu.log 2>&1; echo "2gpu tests rc=$?"; tail -4 gpurun_out/r2k_tests_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2k_bench_2gpu.json 2> gpurun_out/r2k_bench_2gpu.err; echo "bench2 rc=$?"; tail -c 600 gpurun_out/r2k_bench_2gpu.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2k_bench_2gpu.json'))
print('N=2 value %.1fM ms/step %.4f | eval %.1fM users/s (%.3f ms) e2e %.1f ms | sharded %.1fM inter/s %.3f ms/step'%(b['value']/1e6,b['ms_per_step'],b['eval']['value']/1e6,b['eval']['seconds']*1e3,b['eval']['e2e']['seconds']*1e3,b['sharded_train']['value']/1e6,b['sharded_train']['ms_per_step']))
PY
