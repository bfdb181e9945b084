# usage: bash tools/gpu_full.sh <tag> -- GPU tests, bench (1 GPU), ncu launch list, ncu --set full of the steady-state kernels
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 1000 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_[a-z]' -c 6000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-timing --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "kernel_timing/" \
    -k regex:'k_raster_bwd|k_raster_fwd|k_bwd_params|k_adam_rest|k_sort_tiles|k_scatter_tiles|k_project_sh|k_integrate_tma|k_raycast' \
    -c 40 -o gpurun_out/stage_$TAG python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --timing-reps 1 > gpurun_out/ncu_stage_$TAG.log 2>&1
tail -3 gpurun_out/ncu_stage_$TAG.log | cut -c 1-300
ncu -i gpurun_out/stage_$TAG.ncu-rep --page raw --csv > gpurun_out/stage_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
