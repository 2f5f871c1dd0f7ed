# precision study + launch list + ncu full capture of one step + sanitizer pass
mkdir -p gpurun_out
timeout 900 python tools/precision_study.py --iters 50 --res 24 > gpurun_out/r2_precision_study.json 2> gpurun_out/r2_precision_study.err; echo "precision rc=$?" > gpurun_out/r2d_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_c3_x3.csv python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu-baseline --quick > gpurun_out/r2_launches.log 2>&1; echo "launchlist rc=$?" >> gpurun_out/r2d_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'raymarch|splat_wavg|conv3x3_halo|conv_first|smooth3|gram|avgpool|adam_iterate' --launch-skip 60 --launch-count 40 -f -o gpurun_out/r2_ncu_full_c3_x3 python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline --quick > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/r2d_summary.txt
bash tools/gpu_sanitize.sh > gpurun_out/r2_sanitize.log 2>&1; echo "sanitize rc=$?" >> gpurun_out/r2d_summary.txt
cat gpurun_out/r2d_summary.txt gpurun_out/sanitizer_summary.txt; ls -la gpurun_out/*.ncu-rep
