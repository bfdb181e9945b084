# usage: bash tools/gpu_ref_gpu.sh <tag> -- GPU-vs-GPU denominators: reference InfiniTAM CUDA engine (sm_100a) vs ours, reference gsplat-kernel loop vs ours
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python tools/time_reference_itm_cuda.py --frames 400 --timed 100 > gpurun_out/ref_itm_cuda_$TAG.json 2> gpurun_out/ref_itm_cuda_$TAG.err; tail -3 gpurun_out/ref_itm_cuda_$TAG.err; cat gpurun_out/ref_itm_cuda_$TAG.json
timeout 900 python tools/ref_loop.py --frames 81 > gpurun_out/ref_loop_timed_$TAG.json 2> gpurun_out/ref_loop_timed_$TAG.err; tail -3 gpurun_out/ref_loop_timed_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/ref_loop_timed_$TAG.json").read().strip().splitlines()[-1])
for k in ("engine","reference_kernels"):
    print(k, {x: d[k][x] for x in ("psnr_db","loop_seconds","frames_per_sec","last_cycle_ms")}, d[k]["gaussians_after_each_cycle"][-1])
print("psnr_vs_reference_db", d["psnr_vs_reference_db"])
PY
