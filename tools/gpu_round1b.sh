mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_tc_gpu.py -x -q --timeout=180 > gpurun_out/pytest_tc.log 2>&1; echo "pytest tc rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?" >> gpurun_out/summary.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --workload C3 --steps 10 --warmup 3 > gpurun_out/bench_c3_bf16.json 2> gpurun_out/bench_c3_bf16.err; echo "bench c3 rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload C2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_bf16.json 2> gpurun_out/bench_c2_bf16.err; echo "bench c2 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -25 gpurun_out/pytest_tc.log; tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; head -c 2500 gpurun_out/bench_c3_bf16.json; tail -3 gpurun_out/bench_c3_bf16.err
