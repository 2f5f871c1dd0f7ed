python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'raymarch|splat_wavg|conv3x3_tc|conv_first' --launch-skip 14 --launch-count 14 -o gpurun_out/prof_r1_hot python bench.py --workload C3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep; tail -5 gpurun_out/ncu_full.log
