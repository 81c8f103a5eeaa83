mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit_eval.py -m gpu -q -x -k "tc or topk" > gpurun_out/r2az_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2az_tests.log
for TR in 1 0; do
MFB_TC_TILE_RADIUS=$TR EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2az_eval_tr$TR.log 2>&1; echo "tile_radius=$TR: $(grep 'MFB_TC=1\|candidate' gpurun_out/r2az_eval_tr$TR.log | tr '\n' ' ')"
MFB_TC_TILE_RADIUS=$TR timeout 300 python tools/eval_shard_probe.py 2>&1 | head -1
done
