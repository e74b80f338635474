mkdir -p gpurun_out; : > gpurun_out/m.log
for m in 2 0; do
  BT_TC4_MODE=$m timeout 120 python scripts/tc4_check.py 10 0.5 10 4000000 2>&1 | grep "WORST\|TIMING\|rror" | sed "s/^/mode $m: /" >> gpurun_out/m.log
done
BT_TC4_MODE=2 BT_TC_DEBUG=1 timeout 100 python scripts/tc4_time.py 4000000 > gpurun_out/tl_m2.log 2>&1
cat gpurun_out/m.log
