python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --workload C3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload C2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -8 gpurun_out/pytest_gpu.log; python - <<'PY'
import json
for f in ('gpurun_out/bench_c3.json','gpurun_out/bench_c2.json'):
    try:
        d=json.load(open(f)); print(f, 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['roofline']['kernel'], round(d['roofline']['frac'],3)); print(d['kernel_table_ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/bench_c3.err
