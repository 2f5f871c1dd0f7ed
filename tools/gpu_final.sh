# round-2 final evidence run (one GPU): full -m gpu suite, the default bench line, the reference arm, launch list, smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2_final_summary.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_c3_final.json 2> gpurun_out/r2_bench_c3_final.err; echo "bench rc=$?" >> gpurun_out/r2_final_summary.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_c3_reference_arm.json 2> gpurun_out/r2_bench_ref.err; echo "reference arm rc=$?" >> gpurun_out/r2_final_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_c3_x3.csv python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu-baseline --quick > gpurun_out/r2_launches.log 2>&1; echo "launchlist rc=$?" >> gpurun_out/r2_final_summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_final_summary.txt
cat gpurun_out/r2_final_summary.txt; tail -2 gpurun_out/r2_pytest_gpu_final.log; tail -1 gpurun_out/r2_smoke.log
