# usage: bash tools/gpu_final.sh <tag> -- what the driver runs at round end (GPU suite, smoke, bench, reference arm) + config 3 + the ncu evidence
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/suite_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/smoke_$TAG.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 400 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; cut -c1-400 gpurun_out/bench_reference_$TAG.json
timeout 600 python bench.py --config 3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_config3_$TAG.json 2> gpurun_out/bench_config3_$TAG.err; tail -c 300 gpurun_out/bench_config3_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("fps %.1f e2e %.1f full_run %.1f psnr %.3f launches %d cpu %s" % (d["value"], d["e2e"]["value"], d["config"]["full_run"]["fps"], d["config"]["quality"]["psnr_db"], d["gpu_launches"], d["cpu_baseline"]))
r=d["roofline"]; print("roofline", r["frac"], r["avg_launch_us"], r.get("issue_bound"), r.get("pairs", {}).get("issue_slots_per_tested_pair"))
print("  " + ", ".join("%s %.0f" % (k.split("(")[0], v) for k, v in r["kernels_us"].items()))
try:
    t=json.load(open("gpurun_out/bench_config3_$TAG.json"))
    print("config 3: fps %.1f e2e %.1f tracking %s" % (t["value"], t["e2e"]["value"], t["config"]["tracking"]))
except Exception as e:
    print("config 3:", e)
PY
GSB_PROFILE_WINDOW=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 5 --no-cpu-baseline --no-kernel-timing --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
tail -1 gpurun_out/ncu_launch_$TAG.log | cut -c 1-200
timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "kernel_timing/" \
    -k regex:'k_raster_bwd|k_raster_fwd|k_bwd_params|k_adam_rest|k_sort_tiles|k_scatter_tiles|k_project_sh|k_integrate_tma|k_raycast' \
    -c 30 -o /tmp/stage_$TAG python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --timing-reps 1 > gpurun_out/ncu_stage_$TAG.log 2>&1
tail -1 gpurun_out/ncu_stage_$TAG.log | cut -c 1-200
ncu -i /tmp/stage_$TAG.ncu-rep --page raw --csv > gpurun_out/stage_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
