mkdir -p gpurun_out
timeout 300 python tools/eval_shard_probe.py 2>&1 | tee gpurun_out/r2ao_eval_shard_probe.txt
MFB_TC_THR_TPU4=65536 timeout 300 python tools/eval_shard_probe.py 2>&1 | tee gpurun_out/r2ao_eval_shard_probe_tpu4.txt
EVAL_U=17312 EVAL_TC_ONLY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_tc_|k_topk" -c 40 --csv --log-file gpurun_out/r2ao_launches_u17312.csv python tools/eval_bench.py > gpurun_out/r2ao_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2ao_launches_u17312.csv | cut -c1-140
