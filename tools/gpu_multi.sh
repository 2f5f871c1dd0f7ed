# 2-GPU checks: NCCL tests (views sharded + frames sharded == 1 GPU) and the N=2 bench line
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py tests/test_widen_zz_multigpu_multinet.py -m gpu -q --timeout=300 > gpurun_out/pytest_multigpu.log 2>&1; echo "pytest multigpu rc=$?"; tail -3 gpurun_out/pytest_multigpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; echo "bench n2 rc=$?"
head -c 700 gpurun_out/bench_c3_n2.json; tail -2 gpurun_out/bench_c3_n2.err
