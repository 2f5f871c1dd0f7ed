# final evidence of a round: GPU tests, default bench line (with the CPU baseline), the reference arm, launch list and
# ncu --set full of the hot kernels -> gpurun_out/  (1 GPU)
TAG=${1:-r1_final2}
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench ref rc=$?"
timeout 600 python bench.py --workload C2 --no-cpu-baseline > gpurun_out/bench_c2_$TAG.json 2> /dev/null; echo "bench c2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_$TAG.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_halo|conv3x3_tc_persist|raymarch_rot|splat_wavg|smooth3|gram_tc|conv_first|avgpool|adam_iterate' --launch-skip 70 --launch-count 36 -o gpurun_out/prof_$TAG python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/prof_$TAG.ncu-rep
head -c 600 gpurun_out/bench_$TAG.json; echo; cat gpurun_out/bench_ref_$TAG.json | head -c 400
