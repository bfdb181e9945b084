# usage: bash tools/gpu_cxx.sh -- the C++ host layer tests (gsplat wrappers over the C ABI; InfiniTAM facade driver)
mkdir -p gpurun_out
F="tests/test_cxx_shim_gpu.py"; [ -f tests/test_cxx_itm_gpu.py ] && F="$F tests/test_cxx_itm_gpu.py"
timeout 1200 python -m pytest $F -m gpu -q 2>&1 | tail -40
