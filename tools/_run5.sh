mkdir -p gpurun_out
for V in libmfb200 var_st3_scr16 var_st3_scr8 var_st4_scr12 var_st4_scr8; do
  for CLU in 1 2; do
    MFB_LIB_PATH=recommendation_gans_b200/lib/$V.so MFB_TC_CLUSTER=$CLU EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2e_eval_${V}_cl$CLU.log 2>&1; echo "$V cluster=$CLU rc=$? $(grep 'MFB_TC=1' gpurun_out/r2e_eval_${V}_cl$CLU.log)"
  done
done
for V in libmfb200 var_st4_scr8; do
  for CLU in 1 2; do
    MFB_TC_DBG=1 MFB_LIB_PATH=recommendation_gans_b200/lib/$V.so MFB_TC_CLUSTER=$CLU EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2e_evaldbg1_${V}_cl$CLU.log 2>&1; echo "DBG1 $V cluster=$CLU rc=$? $(grep 'MFB_TC=1' gpurun_out/r2e_evaldbg1_${V}_cl$CLU.log)"
  done
done
# skewed (trained-shape) tables with the best candidates
for V in libmfb200 var_st4_scr8; do
  EVAL_SKEW=2 MFB_LIB_PATH=recommendation_gans_b200/lib/$V.so MFB_TC_CLUSTER=2 EVAL_TC_ONLY=1 timeout 200 python tools/eval_bench.py > gpurun_out/r2e_evalskew_${V}.log 2>&1; echo "SKEW $V rc=$? $(grep 'MFB_TC=1\|candidate' gpurun_out/r2e_evalskew_${V}.log | tr '\n' ' ')"
done
B="python bench.py --no-twin --no-sharded --no-cpu-baseline"
for K in 20 494; do
  for R in 0 8; do
    MFB_CHUNK_RAMP=$R timeout 300 $B --steps $K --warmup 5 > gpurun_out/r2e_b_ramp${R}_$K.json 2>/dev/null
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2e_b_*.json')):
    try:
        b=json.load(open(f))
        print(f.split('r2e_b_')[1], 'ms/step %.4f [%.4f..%.4f] e2e %.1fM upd %.1fus k_us %s'%(b['ms_per_step'], b['timing']['ms_per_step_min'], b['timing']['ms_per_step_max'], b['e2e']['value']/1e6, b['roofline']['us_per_launch'], {k:round(v,1) for k,v in b['kernel_us_per_step'].items()}))
    except Exception as e: print(f, 'ERR', e)
PY
