for DBG in 1 2 16 18; do
MFB_TC_DBG=$DBG EVAL_TC_ONLY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_tc_gemm" -c 4 --csv --log-file gpurun_out/r2n_dbg$DBG.csv python tools/eval_bench.py > /dev/null 2>&1
echo "dbg=$DBG"; python tools/ncu_summary.py gpurun_out/r2n_dbg$DBG.csv | grep k_tc_gemm
done
