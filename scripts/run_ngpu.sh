# usage: run_ngpu.sh N  -- bench on N GPUs of one box (torchrun), line kept under gpurun_out/benchN.log
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench$N.log 2>gpurun_out/bench$N.err; echo "rc=$?" >> gpurun_out/bench$N.err
wc -l gpurun_out/bench$N.log; head -c 200 gpurun_out/bench$N.log; tail -1 gpurun_out/bench$N.err
