# First GPU pass over the widening rows: parity through the C-ABI, then timings and a launch list.
#   gpurun --timeout 1500 -- 'bash tools/gpu_widen.sh'
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -k "widen" > gpurun_out/pytest_widen.log 2>&1; echo "pytest widen rc=$?" > gpurun_out/widen_summary.txt
timeout 600 python tools/wbench.py > gpurun_out/wbench.jsonl 2> gpurun_out/wbench.err; echo "wbench rc=$?" >> gpurun_out/widen_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_widen.csv \
    python tools/wbench.py --reps 1 --only g2p,rk4,inception > /dev/null 2>&1; echo "ncu rc=$?" >> gpurun_out/widen_summary.txt
timeout 900 python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?" >> gpurun_out/widen_summary.txt
cat gpurun_out/widen_summary.txt; tail -5 gpurun_out/pytest_widen.log; cat gpurun_out/wbench.jsonl; tail -3 gpurun_out/wbench.err
