mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/gputests.log 2>&1
echo "tests rc=$?" >> gpurun_out/gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err
tail -4 gpurun_out/gputests.log; tail -1 gpurun_out/smoke.log; head -c 400 gpurun_out/bench.log
