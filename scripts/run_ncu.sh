# ncu evidence for the round: (1) launch list of the bench command, (2) full capture of the forward-map kernels
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python scripts/prof_forward.py > gpurun_out/prof_plain.log 2>&1 || exit 2
ncu --set full --clock-control none --import-source on -k regex:"mlp_tc4_kernel|mlp_pre_mma_kernel" -c 2 -o gpurun_out/r02_forward_map \
    python scripts/prof_forward.py > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:"mtm_nl_kernel|mtm_gen_kernel|pmala_accept_kernel|pmala_propose_kernel|mtm_select_kernel" -c 10 -o gpurun_out/r02_step_kernels \
    python bench.py --steps 2 --warmup 0 --size 512 --no-extras --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_full.log
