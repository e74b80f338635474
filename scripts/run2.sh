mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/gputests.log 2>&1
echo "tests rc=$?" >> gpurun_out/gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err
tail -5 gpurun_out/gputests.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_ref.log
