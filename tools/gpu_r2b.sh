mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tma_gpu.py tests/test_tc_x3_gpu.py -q --timeout=600 -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2b_summary.txt
timeout 600 python bench.py --workload C3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_c3_bf16.json 2> gpurun_out/r2b_bench_c3_bf16.err; echo "bench rc=$?" >> gpurun_out/r2b_summary.txt
cat gpurun_out/r2b_summary.txt; tail -30 gpurun_out/r2b_pytest.log | cut -c1-300
python - <<'PY'
import json
for f in ('gpurun_out/r2b_bench_c3_bf16.json',):
    try:
        d=json.load(open(f)); print(f, 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1)); print(d.get('kernel_table_ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/r2b_bench_c3_bf16.err
