# usage: bash tools/gpu_r2.sh <tag> [pytest args...] -- GPU suite (or the given tests) + the driver-shaped bench
TAG=${1:-x}; shift
mkdir -p gpurun_out
if [ "$#" -gt 0 ]; then
  timeout 1500 python -m pytest "$@" -m gpu -q -x 2>&1 | tail -15
else
  timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
fi
timeout 900 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 1500 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("fps %.1f e2e %.1f launches %d" % (d["value"], d["e2e"]["value"], d["gpu_launches"]))
print("quality", d["config"]["quality"])
print("full_run", d["config"]["full_run"])
print("each", d["config"]["ms_per_step_each"])
print({k: d["config"][k] for k in ("gaussians","allocated_blocks","last_isects","last_visible","overflow_flags","keyframes","opt_cameras") if k in d["config"]})
r=d["roofline"]
if r:
    print("roofline", {k: r[k] for k in ("kernel","achieved","frac","avg_launch_us","traffic") if k in r})
    for k,v in r.get("kernels_us",{}).items(): print("  %-40s %8.1f us" % (k, v))
    ti=r.get("tsdf_integrate") or {}
    print("integrate", {k: ti.get(k) for k in ("achieved","frac","avg_launch_us","fresh_frame_stages_us")})
print("cpu", d["cpu_baseline"])
PY
