# round 2, call A: the whole -m gpu suite (no -x: every failure is listed), then C3 bench lines for bf16 and bf16x3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2a_summary.txt
timeout 600 python bench.py --workload C3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c3_bf16.json 2> gpurun_out/r2a_bench_c3_bf16.err; echo "bench bf16 rc=$?" >> gpurun_out/r2a_summary.txt
timeout 600 python bench.py --workload C3 --steps 20 --warmup 3 --no-cpu-baseline --conv-math bf16x3 > gpurun_out/r2a_bench_c3_x3.json 2> gpurun_out/r2a_bench_c3_x3.err; echo "bench x3 rc=$?" >> gpurun_out/r2a_summary.txt
cat gpurun_out/r2a_summary.txt; tail -25 gpurun_out/r2a_pytest_gpu.log
python - <<'PY'
import json
for f in ('gpurun_out/r2a_bench_c3_bf16.json','gpurun_out/r2a_bench_c3_x3.json'):
    try:
        d=json.load(open(f)); print(f, 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1)); print(d.get('kernel_table_ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/r2a_bench_c3_x3.err
