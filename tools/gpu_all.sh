# usage: bash tools/gpu_all.sh <tag> -- the whole GPU suite + bench without the CPU leg
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("fps %.1f e2e %.1f psnr %.3f launches %d" % (d["value"], d["e2e"]["value"], d["config"]["quality"]["psnr_db"], d["gpu_launches"]))
for k,v in d["roofline"]["kernels_us"].items(): print("  %-40s %8.1f us" % (k, v))
PY
