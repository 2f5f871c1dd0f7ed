mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload tiny --steps 5 --warmup 3 --conv-math fp32 --no-cpu-baseline > gpurun_out/bench_tiny_fp32.json 2> gpurun_out/bench_tiny_fp32.err; echo "bench tiny rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --workload C3 --steps 5 --warmup 3 --conv-math fp32 > gpurun_out/bench_c3_fp32.json 2> gpurun_out/bench_c3_fp32.err; echo "bench c3 rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload C2 --steps 10 --warmup 3 --conv-math fp32 --no-cpu-baseline > gpurun_out/bench_c2_fp32.json 2> gpurun_out/bench_c2_fp32.err; echo "bench c2 rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1_fp32.csv python bench.py --workload C3 --steps 1 --warmup 1 --no-cpu-baseline --conv-math fp32 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_c3_fp32.json | head -c 3000
