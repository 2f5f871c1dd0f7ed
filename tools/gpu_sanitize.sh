# compute-sanitizer pass over the C-ABI kernels on a B200 (memcheck, then racecheck on the shared-memory kernels):
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh'
# Runs the kernel-level parity tests (small shapes, every SIMT entry point) under the sanitizer; the tcgen05 kernels
# are covered by memcheck only (racecheck does not model TMA / mbarrier traffic).  Logs land in gpurun_out/.
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
mkdir -p gpurun_out
SEL='tests/test_kernel_parity.py tests/test_widen_graphnet_kernels.py tests/test_widen_resim.py tests/test_widen_style_mask.py tests/test_r2_boundary.py tests/test_tc_x3_gpu.py tests/test_tma_gpu.py'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_memcheck.log \
    python -m pytest $SEL -m gpu -q -x --timeout=900 > gpurun_out/sanitizer_memcheck_pytest.log 2>&1
echo "memcheck rc=$?" > gpurun_out/sanitizer_summary.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/sanitizer_racecheck.log \
    python -m pytest tests/test_kernel_parity.py tests/test_widen_graphnet_kernels.py -m gpu -q -x --timeout=900 \
    -k "conv or gram or pressure or smooth or image_max or lrn" > gpurun_out/sanitizer_racecheck_pytest.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_summary.txt
cat gpurun_out/sanitizer_summary.txt; tail -3 gpurun_out/sanitizer_memcheck.log gpurun_out/sanitizer_racecheck.log
