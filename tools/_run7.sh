mkdir -p gpurun_out; rm -f gpurun_out/config_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r2g_tests.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2g_smoke.log
