mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?" > gpurun_out/r2c_summary.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_summary.txt
cat gpurun_out/r2c_summary.txt; tail -5 gpurun_out/r2c_pytest_gpu.log | cut -c1-300; tail -5 gpurun_out/r2c_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c_bench.json'))
    print('value', round(d['value'],1), 'bf16', round(d.get('value_bf16',0),1), 'seq', round(d.get('value_sequential',0),1), 'e2e', round(d['e2e']['value'],1), 'e2e_run', d.get('e2e_run',{}).get('value'))
    print(d['kernel_table_ms_per_step'])
    print(json.dumps(d.get('configs'), indent=1)[:3000])
    print(d.get('cpu_baseline'))
except Exception as e: print('ERR', e)
PY
