mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/gpus.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.log 2>gpurun_out/bench8.err; echo "rc=$?" >> gpurun_out/bench8.err
wc -l gpurun_out/bench8.log; head -c 300 gpurun_out/bench8.log; tail -2 gpurun_out/bench8.err
