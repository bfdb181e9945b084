# diagnostic for order-dependent failures: initcheck on the two tests, then the suspected interaction three times
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool initcheck --print-limit 40 python -m pytest tests/test_gs_edge_gpu.py "tests/test_gs_raw_gpu.py::test_raw_rasteriser_matches_reference" -m gpu -q -x > gpurun_out/initcheck.log 2>&1
echo "initcheck reports: $(grep -c Uninitialized gpurun_out/initcheck.log)"; tail -3 gpurun_out/initcheck.log
for i in 1 2 3; do
  timeout 600 python -m pytest tests/test_checkpoint_gpu.py tests/test_cxx_itm_gpu.py tests/test_cxx_shim_gpu.py tests/test_golden_gpu.py tests/test_gs_edge_gpu.py tests/test_gs_raw_gpu.py -m gpu -q > gpurun_out/flaky_$i.log 2>&1; tail -3 gpurun_out/flaky_$i.log
done
