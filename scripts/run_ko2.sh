mkdir -p gpurun_out; : > gpurun_out/ko2.log
for f in 0 1 16 17 4; do
  BT_TC4_FLAGS=$f timeout 100 python scripts/tc4_time.py 4000000 2>&1 | grep variant | sed "s/^/flags $f: /" >> gpurun_out/ko2.log
done
cut -c1-200 gpurun_out/ko2.log
