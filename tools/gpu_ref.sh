# usage: bash tools/gpu_ref.sh <tag> -- reference-kernel parity tests + per-iteration timing of the reference's gsplat kernels vs ours
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gs_reference_gpu.py -m gpu -q -x 2>&1 | tail -25
timeout 600 python tools/time_reference_gs.py > gpurun_out/ref_gs_timing_$TAG.json 2> gpurun_out/ref_gs_timing_$TAG.err; tail -c 1500 gpurun_out/ref_gs_timing_$TAG.err; cat gpurun_out/ref_gs_timing_$TAG.json
