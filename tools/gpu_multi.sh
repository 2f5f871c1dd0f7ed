# N-GPU checks (gpurun --gpus N): NCCL tests (views sharded + frames sharded == 1 GPU), the N-GPU bench line, C4 frames sharded
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py tests/test_widen_zz_multigpu_multinet.py -m gpu -q -s --timeout=600 -p no:cacheprovider > gpurun_out/r2_pytest_multigpu_n$N.log 2>&1; echo "pytest multigpu rc=$?"; grep -E "passed|failed|2 GPUs vs 1" gpurun_out/r2_pytest_multigpu_n$N.log | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_bench_c3_n$N.json')); print('N=$N value', round(d['value'],1), 'bf16', round(d.get('value_bf16',0),1), 'e2e', round(d['e2e']['value'],1), d['config']['views_per_rank'])
    print(d['kernel_table_ms_per_step'])
except Exception as e: print('ERR', e)
PY
tail -3 gpurun_out/r2_bench_c3_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/c4_sharded.py > gpurun_out/r2_c4_sharded_n$N.json 2> gpurun_out/r2_c4_sharded_n$N.err; echo "c4 rc=$?"; cat gpurun_out/r2_c4_sharded_n$N.json; tail -3 gpurun_out/r2_c4_sharded_n$N.err
