# ncu --set full of one bf16x3 C3 step (kept under 64 MiB: no source import; read here with tools/ncu_summary.py)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'raymarch|splat_wavg|conv3x3_halo|conv3x3_tc_persist|conv_first|smooth3|gram|avgpool|adam_iterate' --launch-skip 60 --launch-count 30 -f -o gpurun_out/r2_ncu_full_c3_x3 python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline --quick > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
