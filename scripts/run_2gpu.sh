mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.log
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/multi.log 2>&1; echo "rc=$?" >> gpurun_out/multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.log 2>gpurun_out/bench2.err; echo "rc=$?" >> gpurun_out/bench2.err
tail -3 gpurun_out/multi.log; head -c 300 gpurun_out/bench2.log; tail -2 gpurun_out/bench2.err
