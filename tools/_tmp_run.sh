for V in libmfb200 var_rs8 var_rs2 libmfb200 var_rs8; do
  MFB_LIB_PATH=recommendation_gans_b200/lib/$V.so timeout 300 python tools/eval_shard_probe.py 2>&1 | head -1 | sed "s/^/$V: /"
done
