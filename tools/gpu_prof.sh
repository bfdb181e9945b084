# usage: bash tools/gpu_prof.sh <tag>   -- tests, bench, ncu launch list + full capture of the top kernels (1 GPU)
TAG=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 1500 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_[a-z]' -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-kernel-timing > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_raster_bwd|k_bwd_params_adam|k_raster_fwd|k_project_sh' -s 160 -c 8 -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-kernel-timing > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
