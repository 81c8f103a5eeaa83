mkdir -p gpurun_out
B="python bench.py --no-twin --no-sharded --no-cpu-baseline"
for IT in uniform zipf; do
for V in libmfb200 var_nopf; do
  MFB_LIB_PATH=recommendation_gans_b200/lib/$V.so timeout 300 $B --items $IT --steps 494 --warmup 5 > gpurun_out/r2i_b_${V}_$IT.json 2>/dev/null
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i_b_*.json')):
    try:
        b=json.load(open(f))
        print(f.split('r2i_b_')[1], 'ms/step %.4f upd %.1fus fwd %.1f'%(b['ms_per_step'], b['roofline']['us_per_launch'], b['kernel_us_per_step']['forward']))
    except Exception as e: print(f, 'ERR', e)
PY
