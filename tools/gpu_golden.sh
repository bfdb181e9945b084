# generate the GS golden fixture from the reference kernels, then run the reference + golden GPU tests
mkdir -p gpurun_out
python tests/golden/make_golden_gs.py gpurun_out/gs_ref_golden.npz
cp gpurun_out/gs_ref_golden.npz tests/golden/gs_ref_golden.npz
timeout 900 python -m pytest tests/test_gs_reference_gpu.py tests/test_golden_gpu.py -m gpu -q 2>&1 | tail -25
timeout 300 python -m pytest tests/test_golden_cpu.py -q 2>&1 | tail -5
