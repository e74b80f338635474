mkdir -p gpurun_out
timeout 300 python scripts/tc4_check.py 10 0.5 10 4000000 > gpurun_out/tc4_check.log 2>&1
timeout 200 env BT_TC_DEBUG=1 python scripts/tc4_time.py 4000000 > gpurun_out/tc4_timeline.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1
timeout 400 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err
tail -3 gpurun_out/tc4_check.log; tail -3 gpurun_out/gputests.log; cat gpurun_out/bench.log
