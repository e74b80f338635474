mkdir -p gpurun_out
timeout 200 python scripts/tc4_check.py 10 0.5 10 4000000 2>&1 | grep -v Warning | tail -4 > gpurun_out/t.log
cat gpurun_out/t.log
