# final evidence for profiles/: launch list of bench steps + ncu --set full of the hot kernels (1 GPU)
TAG=${1:-r1_final}
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_$TAG.csv
timeout 900 ncu --set full --clock-control none -k regex:'conv3x3_halo|conv3x3_tc_persist|raymarch_rot|splat_wavg|smooth3|gram_tc|conv_first' --launch-skip 62 --launch-count 30 -o gpurun_out/prof_$TAG python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/prof_$TAG.ncu-rep
