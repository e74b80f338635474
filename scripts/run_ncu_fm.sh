mkdir -p gpurun_out
python scripts/prof_forward.py > gpurun_out/prof_plain.log 2>&1 || exit 2
ncu --set full --clock-control none --import-source on -k regex:"mlp_tc4_kernel|mlp_pre_mma_kernel" -c 2 -o gpurun_out/r02_forward_map \
    python scripts/prof_forward.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
