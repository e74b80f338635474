# A/B timing: arguments are lib:mode pairs (libraries under gpurun_ab/, BT_TC4_MODE), same box, interleaved
mkdir -p gpurun_out; : > gpurun_out/ab.log
for rep in 1 2; do for v in "$@"; do
  lib=${v%%:*}; mode=${v##*:}
  BT_TC4_MODE=$mode BT_LIB_PATH=$PWD/gpurun_ab/lib$lib.so timeout 100 python scripts/tc4_time.py 4000000 2>&1 | grep "variant\|rror" | sed "s/^/$v: /" >> gpurun_out/ab.log
done; done
for v in "$@"; do
  lib=${v%%:*}; mode=${v##*:}
  BT_TC4_MODE=$mode BT_LIB_PATH=$PWD/gpurun_ab/lib$lib.so timeout 100 python scripts/tc4_time.py 4000000 jac 2>&1 | grep "variant\|rror" | sed "s/^/$v: /" >> gpurun_out/ab.log
done
cat gpurun_out/ab.log | cut -c1-200
