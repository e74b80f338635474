# correctness of the last variant, then interleaved A/B timing (same box)
mkdir -p gpurun_out; : > gpurun_out/ab.log
last="${@: -1}"
BT_LIB_PATH=$PWD/gpurun_ab/lib$last.so timeout 200 python scripts/tc4_check.py 10 0.5 10 0 2>&1 | grep "WORST\|rror" | sed "s/^/$last check: /" >> gpurun_out/ab.log
for rep in 1 2; do for v in "$@"; do
  BT_LIB_PATH=$PWD/gpurun_ab/lib$v.so timeout 100 python scripts/tc4_time.py 4000000 2>&1 | grep "variant\|rror" | sed "s/^/$v: /" >> gpurun_out/ab.log
done; done
for v in "$@"; do
  BT_LIB_PATH=$PWD/gpurun_ab/lib$v.so timeout 100 python scripts/tc4_time.py 4000000 jac 2>&1 | grep "variant\|rror" | sed "s/^/$v: /" >> gpurun_out/ab.log
done
cut -c1-200 gpurun_out/ab.log
