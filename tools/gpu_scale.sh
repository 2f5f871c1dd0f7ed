# the driver's scaling leg for one N: python -m torch.distributed.run ... bench.py --gpus N (default steps)
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_bench_c3_n$N.json')); print('N=$N value', round(d['value'],1), 'bf16', round(d.get('value_bf16',0),1), 'e2e', round(d['e2e']['value'],1), d['config']['views_per_rank'])
    print(d['kernel_table_ms_per_step'])
except Exception as e: print('ERR', e)
PY
tail -3 gpurun_out/r2_bench_c3_n$N.err
