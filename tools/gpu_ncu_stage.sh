# full ncu capture of the steady-state kernels (end of the bench, N ~ 275k Gaussians), 1 GPU
TAG=${1:-x}
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "kernel_timing/" \
    -k regex:'k_raster_bwd|k_raster_fwd|k_bwd_params|k_adam_rest|k_sort_tiles|k_scatter_tiles|k_project_sh|k_integrate_tma|k_raycast' \
    -c 40 -o gpurun_out/stage_$TAG python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --timing-reps 1 > gpurun_out/ncu_stage_$TAG.log 2>&1
tail -3 gpurun_out/ncu_stage_$TAG.log | cut -c 1-300
