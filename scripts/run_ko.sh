# knock-out timings of the column-major pair kernel (results invalid, timing only)
mkdir -p gpurun_out
# (build first, here: BT_NVCC_DEFS=-DBT_KNOCKOUTS python beetroots_b200/build.py --force)
: > gpurun_out/ko.log
for f in 0 1 4 8 16 17 21 32 96; do
  BT_TC4_FLAGS=$f timeout 100 python scripts/tc4_time.py 4000000 2>&1 | sed "s/^/flags $f: /" >> gpurun_out/ko.log
done
BT_TC4_FLAGS=0 timeout 100 python scripts/tc4_time.py 4000000 jac 2>&1 | sed "s/^/flags 0 jac: /" >> gpurun_out/ko.log
for s in 4 5 6; do BT_TC4_SLOTS=$s timeout 100 python scripts/tc4_time.py 4000000 2>&1 | sed "s/^/slots $s: /" >> gpurun_out/ko.log; done
for n in 144 160 192 256; do BT_TC4_NCMAX=$n timeout 100 python scripts/tc4_time.py 4000000 2>&1 | sed "s/^/ncmax $n: /" >> gpurun_out/ko.log; done
cat gpurun_out/ko.log
