# usage: bash tools/gpu_prof_r2.sh <tag> [launches|full|both] -- ncu launch list of two steady-state steps (frames 1930-1950 of 2000) and/or
# ncu --set full of the steady-state kernels (NVTX range kernel_timing), both on the driver's bench command shape.  The .ncu-rep stays on
# the box (gpurun_out/ is limited to 64 MiB): only the raw CSV page comes back.
TAG=${1:-x}; WHAT=${2:-both}
mkdir -p gpurun_out
if [ "$WHAT" != "full" ]; then
GSB_PROFILE_WINDOW=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 5 --no-cpu-baseline --no-kernel-timing --no-e2e ${BENCH_ARGS} > gpurun_out/ncu_launch_$TAG.log 2>&1
tail -2 gpurun_out/ncu_launch_$TAG.log | cut -c 1-300
fi
if [ "$WHAT" != "launches" ]; then
timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "kernel_timing/" \
    -k regex:'k_raster_bwd|k_raster_fwd|k_bwd_params|k_adam_rest|k_sort_tiles|k_scatter_tiles|k_seg_scan|k_scan_tiles|k_project_sh|k_integrate_tma|k_raycast' \
    -c 40 -o /tmp/stage_$TAG python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --timing-reps 1 ${BENCH_ARGS} > gpurun_out/ncu_stage_$TAG.log 2>&1
tail -3 gpurun_out/ncu_stage_$TAG.log | cut -c 1-300
ncu -i /tmp/stage_$TAG.ncu-rep --page raw --csv > gpurun_out/stage_${TAG}_raw.csv 2>/dev/null
fi
ls -la gpurun_out | tail -6
du -sh gpurun_out
