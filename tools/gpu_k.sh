# usage: [VARIANTS="8 16"] bash tools/gpu_k.sh <tag> [pytest files...] -- kernel iteration loop: the given GPU tests, then the bench (kernel table) per GSB_BWD_LPI variant
TAG=$1; shift
mkdir -p gpurun_out
if [ "$#" -gt 0 ]; then timeout 900 python -m pytest "$@" -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/k_tests_$TAG.log; fi
for V in ${VARIANTS:-8}; do
GSB_BWD_LPI=$V timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_${TAG}_v$V.json 2> gpurun_out/bench_${TAG}_v$V.err; tail -c 300 gpurun_out/bench_${TAG}_v$V.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_v$V.json"))
    print("variant $V: fps %.1f e2e %.1f  full_run %.1f  psnr %.3f  launches %d" % (d["value"], d["e2e"]["value"], d["config"]["full_run"]["fps"], d["config"]["quality"]["psnr_db"], d["gpu_launches"]))
    r=d["roofline"]
    print("  roofline frac %.4f bwd %.1f us  pairs %s" % (r["frac"], r["avg_launch_us"], {k: (round(v, 2) if isinstance(v, float) else v) for k, v in (r.get("pairs") or {}).items() if k != "note"}))
    print("  " + ", ".join("%s %.0f" % (k.split("(")[0], v) for k, v in r["kernels_us"].items()))
except Exception as e:
    print("no bench line:", e)
PY
done
