# launch list of one C3 step (cold-cache, serialised: compare SHARES) -> gpurun_out/launches_<tag>.csv
TAG=${1:-r1}
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_$TAG.csv
