# ncu --set full with source correlation for the tensor-core convolution kernels of one C3 step
TAG=${1:-conv}
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_halo|conv3x3_tc_persist' --launch-skip 10 --launch-count 10 -o gpurun_out/prof_$TAG python bench.py --workload C3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_$TAG.log
